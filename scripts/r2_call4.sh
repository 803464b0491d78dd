#!/bin/bash
# scripts/r2_call4.sh -- ncu source-level capture of the persistent kernel on the Orpheus-3B shape (one 32-step launch)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2d
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/index.log"; local t0=$(date +%s); timeout -s KILL "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $name ($(( $(date +%s) - t0 )) s)" | tee -a "$OUT/index.log"; }
run ncu_orpheus_pdk 900 ncu --set full --clock-control none --import-source on -k regex:pdk_kernel -s 3 -c 1 -o "$OUT/orpheus_pdk_full" -f python scripts/pdk_timeline.py 120 9999 orpheus 16
tail -n 5 "$OUT/index.log"; ls -la "$OUT"
