#!/bin/bash
# Last GPU call of round 2 (7 GPU-minutes left): the new VAD and server tests, then -- if time remains -- the smoke.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2w}
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/index.log"; local t0=$(date +%s); timeout -s KILL "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $name ($(( $(date +%s) - t0 )) s)" | tee -a "$OUT/index.log"; }
run vad_tests 60 python -m pytest tests/test_vad_gpu.py -q -s -m gpu
tail -n 12 "$OUT/vad_tests.log"
run server_tests 150 python -m pytest tests/test_server_gpu.py -q -s -m gpu
tail -n 25 "$OUT/server_tests.log"
run smoke 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
tail -n 6 "$OUT/smoke.log"
tail -n 4 "$OUT/index.log"
