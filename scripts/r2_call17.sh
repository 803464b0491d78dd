#!/bin/bash
# The closing GPU call of round 2 (4.7 GPU-minutes left): the T5 encoder's first hardware run, then its timing at the flan-t5-large shape.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2x}
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/index.log"; local t0=$(date +%s); timeout -s KILL "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $name ($(( $(date +%s) - t0 )) s)" | tee -a "$OUT/index.log"; }
run t5_tests 80 python -m pytest tests/test_t5_gpu.py -q -s -m gpu
tail -n 25 "$OUT/t5_tests.log"
run t5_timing 80 python scripts/t5_timing.py
tail -n 6 "$OUT/t5_timing.log"
tail -n 3 "$OUT/index.log"
