#!/bin/bash
# scripts/r2_call1.sh -- first gpurun call of round 2: Dia F16 triage, every -m gpu test, the decode-step A/B matrix (Parler-Mini F16, batch 16), the ncu launch list and
# a --set full capture of the decode kernels, the codec lines.  Everything lands under gpurun_out/r2a/.  Each step in its own process under its own timeout.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2a
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/index.log"; local t0=$(date +%s); timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $name ($(( $(date +%s) - t0 )) s)" | tee -a "$OUT/index.log"; }

nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > "$OUT/gpu.csv" 2>&1
nproc > "$OUT/host.txt"; lscpu | head -20 >> "$OUT/host.txt"; cat /sys/fs/cgroup/cpu.max >> "$OUT/host.txt" 2>&1

run dia_triage_default 120 python scripts/dia_f16_triage.py
run dia_triage_unfused 120 env B2TTS_AR_FUSE=0 B2TTS_AR_ATT=plain B2TTS_GEMV_GN=1 python scripts/dia_f16_triage.py
run dia_triage_mma 120 env B2TTS_AR_MMA=1 python scripts/dia_f16_triage.py
run dia_sanitize_mem 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/dia_f16_triage.py
run dia_sanitize_race 400 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/dia_f16_triage.py

run tests_gpu 900 python -m pytest tests -m gpu -q -rxXs -p no:cacheprovider

for cfg in "plain:" "graph:B2TTS_AR_GRAPH=1" "mma:B2TTS_AR_MMA=1" "graph_mma:B2TTS_AR_GRAPH=1 B2TTS_AR_MMA=1"; do
    name=${cfg%%:*}; envs=${cfg#*:}
    run "bench_parler_$name" 600 env $envs python bench.py --workload parler --steps 2 --warmup 1
done
run bench_parler_q8_graph 600 env B2TTS_AR_GRAPH=1 python bench.py --workload parler --parler-dtype q8_0 --steps 2 --warmup 1

run ncu_parler_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file "$OUT/parler_launches.csv" \
    env B2TTS_AR_MMA=1 python bench.py --workload parler --steps 1 --warmup 1
run ncu_parler_full 900 ncu --set full --clock-control none --import-source on -k regex:"gemv_mma|attention_gqa_kernel|layernorm_kernel|argmax_rows" -s 4000 -c 16 -o "$OUT/parler_gemv_att" -f \
    env B2TTS_AR_MMA=1 python bench.py --workload parler --steps 1 --warmup 1

run bench_kokoro 600 python bench.py --no-cpu-baseline
run bench_dac 600 python bench.py --workload dac
run bench_snac 600 python bench.py --workload snac
grep -h '^{' "$OUT"/bench_*.log > "$OUT/bench_lines.jsonl" 2>/dev/null
tail -n 40 "$OUT/index.log"
