#!/bin/bash
# scripts/r2_call13.sh -- configs 4 and 5 as stated (utterance-sharded over the GPUs of the box), on the N GPUs this call was given
set -u
cd "$(dirname "$0")/.."
N=${2:-2}
OUT=gpurun_out/${1:-r2s}
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/index.log"; local t0=$(date +%s); timeout -s KILL "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $name ($(( $(date +%s) - t0 )) s)" | tee -a "$OUT/index.log"; }
run orpheus_q8_n$N 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --workload orpheus --gpus $N --steps 2
run dia_n$N 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --workload dia --gpus $N --steps 2
grep -h '^{' "$OUT"/*.log > "$OUT/bench_lines.jsonl" 2>/dev/null
tail -n 6 "$OUT/index.log"
