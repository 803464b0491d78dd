#!/bin/bash
# scripts/r2_call14.sh -- after the short cross-attention path: the AR suites, Parler timeline and config-3 bench, default line
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2t}
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/index.log"; local t0=$(date +%s); timeout -s KILL "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $name ($(( $(date +%s) - t0 )) s)" | tee -a "$OUT/index.log"; }
run t_ar 900 python -m pytest tests/test_parler_gpu.py tests/test_ar_fullsize_gpu.py tests/test_dia_gpu.py tests/test_orpheus_gpu.py -m gpu -q
run timeline_parler 300 python scripts/pdk_timeline.py 480 450 parler 16
run timeline_dia 400 python scripts/pdk_timeline.py 200 150 dia 2
run bench_parler 300 python bench.py --workload parler --steps 2 --warmup 1
run bench_default 600 python bench.py
grep -h '^{' "$OUT"/bench_*.log > "$OUT/bench_lines.jsonl" 2>/dev/null
tail -n 8 "$OUT/index.log"
