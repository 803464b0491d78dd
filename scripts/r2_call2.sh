#!/bin/bash
# scripts/r2_call2.sh -- first hardware contact of the persistent decode kernel (each step under its own timeout: a deadlocked cooperative kernel must not hold the box)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2b
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/index.log"; local t0=$(date +%s); timeout -s KILL "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $name ($(( $(date +%s) - t0 )) s)" | tee -a "$OUT/index.log"; }
run pdk_check 120 python scripts/pdk_gpu_check.py
run pdk_check_f32kv 120 env B2TTS_KV=f32 python scripts/pdk_gpu_check.py
run pdk_check_grid 120 env B2TTS_PDK_GRID=37 B2TTS_AR_EXIT_EVERY=2 python scripts/pdk_gpu_check.py
run pdk_memcheck 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/pdk_gpu_check.py
run bench_parler_pdk 300 python bench.py --workload parler --steps 2 --warmup 1
run bench_parler_pdk_f32kv 300 env B2TTS_KV=f32 python bench.py --workload parler --steps 2 --warmup 1
run bench_parler_ops 300 env B2TTS_AR_PDK=0 python bench.py --workload parler --steps 2 --warmup 1
grep -h '^{' "$OUT"/bench_*.log > "$OUT/bench_lines.jsonl" 2>/dev/null
tail -n 20 "$OUT/index.log"; cat "$OUT/pdk_check.log"
