"""Turn gpurun_out/ ncu artefacts into the committed text summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv profiles/r1_launches_summary.txt [steps_in_capture]
    python profiles/summarize.py full     gpurun_out/prof_gemm_r1.ncu-rep profiles/r1_conv_gemm_ncu_full.txt
"""
import collections
import csv
import subprocess
import sys


def launches(src, dst, steps=4):
    lines = [l for l in open(src) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    mi, ui = hdr.index("Metric Name"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    dram = collections.defaultdict(lambda: [0, 0.0])   # kernel -> [launches, DRAM bytes read + written]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    n = 0
    for row in r:
        try:
            v = float(row[vi].replace(",", ""))
        except ValueError:
            continue
        if row[mi].startswith("dram__bytes"):
            nm = row[ki].split("(")[0].replace("void ", "").replace("unnamed>::", "")
            dram[nm][1] += v * scale.get(row[ui], 1.0)
            if row[mi] == "dram__bytes_read.sum":
                dram[nm][0] += 1
            continue
        if row[mi] != "gpu__time_duration.sum":
            continue
        name = row[ki].split("(")[0].replace("void ", "").replace("unnamed>::", "")
        agg[name][0] += 1
        agg[name][1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 1 --warmup 3 --no-cpu-baseline\n")
        f.write(f"# {n} launches over {steps} forward passes (3 warm-up + 1 timed); cold-cache, serialised: compare SHARES, not absolutes\n")
        f.write(f"# total kernel time {tot / 1e6:.3f} ms  ({tot / 1e6 / steps:.3f} ms per forward, {n // steps} launches per forward)\n")
        f.write(f"{'kernel':40s} {'launches':>9s} {'total_ms':>10s} {'ms/forward':>11s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:40]:40s} {v[0]:9d} {v[1] / 1e6:10.3f} {v[1] / 1e6 / steps:11.3f} {v[1] / tot:7.1%}\n")
        if dram:
            f.write("\n# DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) per launch, same capture\n")
            for k, v in sorted(dram.items(), key=lambda kv: -kv[1][1])[:12]:
                f.write(f"{k[:40]:40s} launches {v[0]:6d}  total {v[1] / 1e9:9.3f} GB  per launch {v[1] / max(v[0], 1) / 1e6:10.3f} MB\n")
    if dram:
        import json
        um = [(k, v) for k, v in dram.items() if "conv_umma" in k]
        nl = sum(v[0] for _, v in um); by = sum(v[1] for _, v in um)
        json.dump({"kernel": "conv_umma_kernel", "launches": nl, "dram_bytes_per_launch": by / max(nl, 1), "source": dst}, open(dst.replace("_launches_summary.txt", "_conv_umma_traffic.json"), "w"))
    print(open(dst).read())


WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic"]


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units, rows = r[0], r[1], r[2:]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on  ({src})\n")
        for w in WANT:
            for i, h in enumerate(hdr):
                if h == w:
                    f.write(f"{w} [{units[i]}]: " + " | ".join(row[i][:70] for row in rows) + "\n")
    print(open(dst).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 4)
    else:
        full(sys.argv[2], sys.argv[3])
