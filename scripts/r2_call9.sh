#!/bin/bash
# scripts/r2_call9.sh -- after "no loads for absent rows 8..15": tests of the three programs, timelines (Orpheus b8, Dia), bench sweeps
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2k2}
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/index.log"; local t0=$(date +%s); timeout -s KILL "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $name ($(( $(date +%s) - t0 )) s)" | tee -a "$OUT/index.log"; }
run t_pdk 900 python -m pytest tests/test_orpheus_gpu.py tests/test_dia_gpu.py tests/test_parler_gpu.py -m gpu -q -k "persistent or parler"
run timeline_orpheus8 400 python scripts/pdk_timeline.py 120 100 orpheus 8
run timeline_dia 400 python scripts/pdk_timeline.py 200 150 dia 2
run bench_orpheus_pdk 600 python bench.py --workload orpheus --orpheus-dtype f16 --steps 2
run bench_dia_pdk 900 python bench.py --workload dia --steps 2
grep -h '^{' "$OUT"/bench_*.log > "$OUT/bench_lines.jsonl" 2>/dev/null
tail -n 12 "$OUT/index.log"
