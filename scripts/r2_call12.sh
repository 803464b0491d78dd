#!/bin/bash
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2n2}
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/index.log"; local t0=$(date +%s); timeout -s KILL "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $name ($(( $(date +%s) - t0 )) s)" | tee -a "$OUT/index.log"; }
run t_orpheus 900 python -m pytest tests/test_orpheus_gpu.py -m gpu -q -s -k "q8_0"
run bench_orpheus_q8_pdk 600 python bench.py --workload orpheus --steps 2
run timeline_q8 400 env B2TTS_TIMELINE_DTYPE=q8_0 python scripts/pdk_timeline.py 120 100 orpheus 8
grep -h '^{' "$OUT"/bench_*.log > "$OUT/bench_lines.jsonl" 2>/dev/null
tail -n 12 "$OUT/index.log"
