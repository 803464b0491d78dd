#!/bin/bash
# scripts/r2_final.sh -- the closing evidence run of round 2 on the final build: the whole -m gpu suite, smoke(), the default bench line and its reference arm, the
# secondary lines (configs 3 / 4 / 5, codecs), the ncu launch list of the default command and of the config-3 step
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2z}
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/index.log"; local t0=$(date +%s); timeout -s KILL "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $name ($(( $(date +%s) - t0 )) s)" | tee -a "$OUT/index.log"; }
run pytest_gpu 2400 python -m pytest tests/ -q -m gpu
run smoke 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run bench_default 600 python bench.py
run bench_reference 600 python bench.py --impl reference --steps 2 --warmup 1
run bench_parler 300 python bench.py --workload parler --steps 2 --warmup 1
run bench_dia 900 python bench.py --workload dia --steps 2
run bench_orpheus_f16 600 python bench.py --workload orpheus --orpheus-dtype f16 --steps 2
run bench_orpheus_q8 600 python bench.py --workload orpheus --steps 2
run bench_dac 300 python bench.py --workload dac
run bench_snac 300 python bench.py --workload snac
run ncu_default_launches 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file "$OUT/default_launches.csv" python bench.py --steps 2 --warmup 1 --no-strong --no-decode-step
grep -h '^{' "$OUT"/bench_*.log > "$OUT/bench_lines.jsonl" 2>/dev/null
tail -n 16 "$OUT/index.log"
