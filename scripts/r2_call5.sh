#!/bin/bash
# scripts/r2_call5.sh -- after the lean GEMV loops: parity, timelines, bench of the persistent kernel on both models
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2e
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/index.log"; local t0=$(date +%s); timeout -s KILL "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $name ($(( $(date +%s) - t0 )) s)" | tee -a "$OUT/index.log"; }
run t_ar 900 python -m pytest tests/test_orpheus_gpu.py tests/test_parler_gpu.py tests/test_ar_fullsize_gpu.py -m gpu -x -q
run timeline_orpheus 400 python scripts/pdk_timeline.py 120 100 orpheus 16
run timeline_parler 300 python scripts/pdk_timeline.py 480 450 parler 16
run bench_orpheus_pdk 600 python bench.py --workload orpheus --orpheus-dtype f16 --steps 2
run bench_parler_pdk 300 python bench.py --workload parler --steps 2 --warmup 1
grep -h '^{' "$OUT"/bench_*.log > "$OUT/bench_lines.jsonl" 2>/dev/null
tail -n 12 "$OUT/index.log"
