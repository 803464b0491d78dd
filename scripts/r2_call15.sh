#!/bin/bash
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2u}
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/index.log"; local t0=$(date +%s); timeout -s KILL "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $name ($(( $(date +%s) - t0 )) s)" | tee -a "$OUT/index.log"; }
run pytest_gpu 2400 python -m pytest tests/ -q -m gpu
run smoke 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
tail -n 4 "$OUT/index.log"
