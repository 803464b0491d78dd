#!/bin/bash
# scripts/r2_call7.sh -- timelines + bench + one ncu source-level capture of the persistent kernel (Orpheus-3B shape)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2h}
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/index.log"; local t0=$(date +%s); timeout -s KILL "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $name ($(( $(date +%s) - t0 )) s)" | tee -a "$OUT/index.log"; }
run t_orpheus 600 python -m pytest tests/test_orpheus_gpu.py -m gpu -x -q -k persistent
run timeline_orpheus 400 python scripts/pdk_timeline.py 120 100 orpheus 16
run timeline_parler 300 python scripts/pdk_timeline.py 480 450 parler 16
run bench_orpheus_pdk 600 python bench.py --workload orpheus --orpheus-dtype f16 --steps 2
run ncu_orpheus_pdk 900 ncu --set full --clock-control none --import-source on -k regex:pdk_kernel -s 3 -c 1 -o "$OUT/orpheus_pdk_full" -f python scripts/pdk_timeline.py 120 9999 orpheus 16
grep -h '^{' "$OUT"/bench_*.log > "$OUT/bench_lines.jsonl" 2>/dev/null
tail -n 12 "$OUT/index.log"
