#!/bin/bash
# scripts/r2_first_gpu_hour.sh -- the first gpurun call of round 2, as ONE command:
#
#     /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/r2_first_gpu_hour.sh'
#
# Row B (the Parler / Dia / Orpheus decode steps, the sampler, CUDA-graph replay, the tensor-core GEMV) has only ever run under the CPU emulation of tests/emu.
# This script runs those paths on the B200, each step in its own process under its own timeout (a fault in one must not cost the rest), and leaves everything a
# reader needs under gpurun_out/r2a/: which -m gpu tests XPASS, the A/B matrix of the environment switches on `bench.py --workload parler`, the ncu launch list of
# one Parler bench step and a `--set full` capture of the two kernels the decode roofline is about.  Nothing here is a bench value: numbers printed under ncu are
# for the kernel's SHARE of a step only.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2a
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/index.log"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $name" | tee -a "$OUT/index.log"; }

nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > "$OUT/gpu.csv" 2>&1

# 0. four seconds each: the small models of all three decode paths against the reference's tokens / logits, default path and two switches
run contact_default 120 python scripts/rowb_first_contact.py
run contact_unfused 120 env B2TTS_AR_FUSE=0 python scripts/rowb_first_contact.py
run contact_mma 120 env B2TTS_AR_MMA=1 python scripts/rowb_first_contact.py

# 1. every -m gpu test, row B included (xfail(strict=False) in child processes): XPASS = works on hardware
run tests_gpu 900 python -m pytest tests -m gpu -q -rxXs -p no:cacheprovider

# 2. the row-B children once more under compute-sanitizer if any of them failed (memcheck on the child command line)
if grep -q "XFAIL" "$OUT/tests_gpu.log"; then
    run sanitizer_orpheus 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_orpheus_gpu.py -m gpu -q -x -rxXs -p no:cacheprovider
    run sanitizer_parler 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parler_gpu.py -m gpu -q -x -rxXs -p no:cacheprovider -k "f32"
    run sanitizer_sampler 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_sampler_gpu.py -m gpu -q -x -rxXs -p no:cacheprovider
fi

# 3. A/B matrix on BASELINE config 3 (Parler-Mini F16, batch 16 x 10 s + DAC decode); 2 timed steps each (a step = 869 decode steps)
for cfg in "plain:" "unfused:B2TTS_AR_FUSE=0" "graph:B2TTS_AR_GRAPH=1" "mma:B2TTS_AR_MMA=1" "graph_mma:B2TTS_AR_GRAPH=1 B2TTS_AR_MMA=1" "graph_mma_plainatt:B2TTS_AR_GRAPH=1 B2TTS_AR_MMA=1 B2TTS_AR_ATT=plain"; do
    name=${cfg%%:*}; envs=${cfg#*:}
    run "bench_parler_$name" 600 env $envs python bench.py --workload parler --steps 2 --warmup 1
done
for dt in q8_0 q5_0; do
    run "bench_parler_${dt}_graph" 600 env B2TTS_AR_GRAPH=1 python bench.py --workload parler --parler-dtype $dt --steps 2 --warmup 1
done

# 4. ncu: launch list of one step (clock control off), then a full capture of the GEMV and attention kernels of the graph-less run (ncu does not see into graph replays by default)
run ncu_parler_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file "$OUT/parler_launches.csv" \
    env B2TTS_AR_MMA=1 python bench.py --workload parler --steps 1 --warmup 1
run ncu_parler_full 900 ncu --set full --clock-control none --import-source on -k regex:"gemv_mma_kernel|attention_gqa_kernel|gemv_rows_h_kernel" -s 3000 -c 12 -o "$OUT/parler_gemv_att" -f \
    env B2TTS_AR_MMA=1 python bench.py --workload parler --steps 1 --warmup 1

# 5. the headline and the codec lines on the same box (regression check of rows A and C against round 1's 7 400-7 570 / 1 524 audio-s/s)
run bench_kokoro 600 python bench.py
run bench_dac 600 python bench.py --workload dac
run bench_snac 600 python bench.py --workload snac
# 6. the reference's CPU arms on this box's host cores (the denominators of the lines above)
run bench_parler_ref 900 python bench.py --workload parler --impl reference
run bench_dac_ref 600 python bench.py --workload dac --impl reference
run bench_snac_ref 600 python bench.py --workload snac --impl reference
grep -h '^{' "$OUT"/bench_*.log > "$OUT/bench_lines.jsonl" 2>/dev/null
tail -n 40 "$OUT/index.log"
