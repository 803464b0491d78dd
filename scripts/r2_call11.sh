#!/bin/bash
# scripts/r2_call11.sh -- Orpheus Q8_0 through the persistent kernel: parity, bench sweep (kernel on / off), timeline
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2m2}
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/index.log"; local t0=$(date +%s); timeout -s KILL "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $name ($(( $(date +%s) - t0 )) s)" | tee -a "$OUT/index.log"; }
run t_orpheus 900 python -m pytest tests/test_orpheus_gpu.py -m gpu -q -s
run bench_orpheus_q8_pdk 600 python bench.py --workload orpheus --steps 2
run bench_orpheus_q8_ops 600 env B2TTS_AR_PDK=0 python bench.py --workload orpheus --steps 1
grep -h '^{' "$OUT"/bench_*.log > "$OUT/bench_lines.jsonl" 2>/dev/null
tail -n 12 "$OUT/index.log"
