#!/bin/bash
# scripts/r2_call3.sh -- Orpheus through the persistent decode kernel: parity tests, the Parler tests again (shared kernel), timeline + bench of the 3B shape in F16
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2c
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/index.log"; local t0=$(date +%s); timeout -s KILL "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $name ($(( $(date +%s) - t0 )) s)" | tee -a "$OUT/index.log"; }
run t_orpheus 600 python -m pytest tests/test_orpheus_gpu.py -m gpu -x -q -s
run t_parler 900 python -m pytest tests/test_parler_gpu.py tests/test_ar_fullsize_gpu.py -m gpu -x -q
run timeline_orpheus 400 python scripts/pdk_timeline.py 120 100 orpheus 16
run bench_orpheus_pdk 600 python bench.py --workload orpheus --orpheus-dtype f16 --steps 2
run bench_orpheus_ops 600 env B2TTS_AR_PDK=0 python bench.py --workload orpheus --orpheus-dtype f16 --steps 2
grep -h '^{' "$OUT"/bench_*.log > "$OUT/bench_lines.jsonl" 2>/dev/null
tail -n 12 "$OUT/index.log"
