#!/bin/bash
# scripts/r2_call10.sh -- split attention (Dia): parity tests, timeline, bench
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2l2}
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/index.log"; local t0=$(date +%s); timeout -s KILL "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $name ($(( $(date +%s) - t0 )) s)" | tee -a "$OUT/index.log"; }
run t_dia 900 python -m pytest tests/test_dia_gpu.py -m gpu -q -s
run timeline_dia 400 python scripts/pdk_timeline.py 200 150 dia 2
run bench_dia_pdk 900 python bench.py --workload dia --steps 2
run bench_dia_nosplit 900 env B2TTS_PDK_TSPLIT=1 python bench.py --workload dia --steps 1
grep -h '^{' "$OUT"/bench_*.log > "$OUT/bench_lines.jsonl" 2>/dev/null
tail -n 12 "$OUT/index.log"
