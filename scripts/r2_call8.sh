#!/bin/bash
# scripts/r2_call8.sh -- Dia through the persistent kernel: parity tests, bench of the 1.6B shape with the kernel on and off; the AR suites again
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2j}
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/index.log"; local t0=$(date +%s); timeout -s KILL "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $name ($(( $(date +%s) - t0 )) s)" | tee -a "$OUT/index.log"; }
run t_dia 600 python -m pytest tests/test_dia_gpu.py -m gpu -q -s
run t_ar 900 python -m pytest tests/test_orpheus_gpu.py tests/test_parler_gpu.py tests/test_ar_fullsize_gpu.py tests/test_ar_graph_gpu.py -m gpu -q
run bench_dia_pdk 900 python bench.py --workload dia --steps 2
run bench_dia_ops 900 env B2TTS_AR_PDK=0 python bench.py --workload dia --steps 1
run bench_parler 300 python bench.py --workload parler --steps 2 --warmup 1
grep -h '^{' "$OUT"/bench_*.log > "$OUT/bench_lines.jsonl" 2>/dev/null
tail -n 12 "$OUT/index.log"
