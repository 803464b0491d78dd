#!/bin/bash
# scripts/r2_profiles.sh -- the evidence run of round 2: persistent-kernel timeline, ncu capture of the persistent kernel, launch list of the config-3 step, secondary bench lines
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2p
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/index.log"; local t0=$(date +%s); timeout -s KILL "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $name ($(( $(date +%s) - t0 )) s)" | tee -a "$OUT/index.log"; }
run timeline 200 python scripts/pdk_timeline.py 480 450
run ncu_pdk_full 600 ncu --set full --clock-control none --import-source on -k regex:pdk_kernel -s 14 -c 1 -o "$OUT/pdk_full" -f python scripts/pdk_timeline.py 480 450
run ncu_parler_launches 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file "$OUT/parler_pdk_launches.csv" python bench.py --workload parler --steps 1 --warmup 1
run bench_parler_pdk 300 python bench.py --workload parler --steps 2 --warmup 1
run bench_parler_ops 300 env B2TTS_AR_PDK=0 python bench.py --workload parler --steps 2 --warmup 1
run bench_dac 300 python bench.py --workload dac
run bench_snac 300 python bench.py --workload snac
run bench_dia 600 python bench.py --workload dia --steps 2
run bench_orpheus 600 python bench.py --workload orpheus --steps 2
grep -h '^{' "$OUT"/bench_*.log > "$OUT/bench_lines.jsonl" 2>/dev/null
tail -n 12 "$OUT/index.log"
