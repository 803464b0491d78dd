#!/usr/bin/env python
"""bench.py -- headline benchmark: audio-seconds/sec for Kokoro-82M fp16, batch 32 synthetic 64-char prompts per B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path (duration pass + generation pass of the whole batch) through the public C-ABI call
(b2tts_kokoro_run_batch) with HOST buffers in and out.  N > 1: launched by torchrun, one rank per GPU, each rank runs its own
batch of 32 independent utterances (weak scaling, no data-path collective: every rank returns its own shard's audio); NCCL is
used only for the barrier and the max-over-ranks of the timings.

  value      : audio-s/s from CUDA-event device time of the forward (events on the library's launching stream)
  e2e.value  : audio-s/s from wall time of the same calls incl. H2D of tokens and D2H of PCM into pinned host memory
  roofline   : the dominant kernel (conv_gemm, tensor-bound): algorithmic FLOPs / CUDA-event time of its launches, live
  cpu_baseline / --impl reference : the UNMODIFIED reference (oracle/_ref/kokoro_ref, built from /root/reference by
               oracle/Makefile) on the host cores, as P worker processes x 4 ggml threads (its own server's
               n-parallelism model, examples/server/server.cpp:225-321), on a bounded sample of the same prompts.

Secondary lines (`--workload dac | snac | parler | orpheus | dia | t5`; never the headline): the codecs, BASELINE configs 3 / 4 / 5 and the T5 conditional-prompt
encoder pass, each with its own `--impl reference` arm.  (`--workload t5` was added after the round's GPU budget ended: its reference arm has run, its CUDA arm
has not -- the T5 numbers in DESIGN section 7.3 come from scripts/t5_timing.py, which makes the same calls.)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 32
N_PHON = 64
CTX_LEN = 128           # GGUF context_length (>= 66 tokens); keeps the reference's worst-case graph reservation small


def host_cpus():
    """What this process may actually use: the affinity mask and the cgroup CPU quota (os.cpu_count() reports the machine, not the lease)."""
    logical = os.cpu_count() or 1
    try:
        aff = len(os.sched_getaffinity(0))
    except Exception:
        aff = logical
    quota = None
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            quota = float(q) / float(per)
    except Exception:
        pass
    model = ""
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip(); break
    except Exception:
        pass
    usable = aff if quota is None else max(1, min(aff, int(quota + 0.5)))
    return {"logical": logical, "affinity": aff, "cgroup_quota_cpus": quota, "usable": usable, "model": model}


def ref_bin(name: str):
    """-> (path, build) of a reference binary for TIMING: the -march=native build of oracle/Makefile's ref_native target when this host's CPU has every ISA
    extension that build was compiled with (it was built on a Sapphire-Rapids-class Xeon; the GPU boxes are Xeon 8562Y+), else the portable x86-64-v3 build
    the golden vectors come from."""
    nat = os.path.join(ROOT, "oracle", "_ref", "native", name)
    macros = os.path.join(ROOT, "oracle", "_ref", "native", "isa_macros.txt")
    if os.path.exists(nat) and os.path.exists(macros):
        need = {"__AVX512F__": "avx512f", "__AVX512BW__": "avx512bw", "__AVX512VL__": "avx512vl", "__AVX512DQ__": "avx512dq", "__AVX512CD__": "avx512cd", "__AVX512VNNI__": "avx512_vnni",
                "__AVX512BF16__": "avx512_bf16", "__AVX512FP16__": "avx512_fp16", "__AVX512VBMI__": "avx512vbmi", "__AVX512VBMI2__": "avx512_vbmi2", "__AVX512IFMA__": "avx512ifma",
                "__AVX512BITALG__": "avx512_bitalg", "__AVX512VPOPCNTDQ__": "avx512_vpopcntdq", "__AMX_TILE__": "amx_tile", "__AMX_INT8__": "amx_int8", "__AMX_BF16__": "amx_bf16",
                "__AVX2__": "avx2", "__FMA__": "fma", "__F16C__": "f16c"}
        try:
            flags = set()
            for ln in open("/proc/cpuinfo"):
                if ln.startswith("flags"):
                    flags = set(ln.split(":", 1)[1].split()); break
            want = [need[m.split()[1]] for m in open(macros) if m.split()[1] in need]
            if all(f in flags for f in want):
                return nat, "-march=native build (AVX-512/VNNI/BF16/AMX host)"
        except Exception:
            pass
    return os.path.join(ROOT, "oracle", "_ref", name), "x86-64-v3 build"


REF_BIN = os.path.join(ROOT, "oracle", "_ref", "kokoro_ref")


def _gguf():
    from tts_cpp_b200.synth import cached_gguf
    return cached_gguf("f16", CTX_LEN, 0)


def _prompts(rank: int):
    from tts_cpp_b200.synth import synthetic_prompts
    return synthetic_prompts(BATCH, N_PHON, seed0=1234 + rank * BATCH)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("bf16_tflops_sustained", 1400.0)), float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML every 10 ms (nvidia-smi -lms fallback)."""

    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.sm, self.mx, self.reasons = [], 0.0, set()
        self.stop_flag = threading.Event()
        self.thread = None
        self.source = None

    def _nvml_loop(self, nv, h):
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.REASONS:
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def _smi_loop(self, proc):
        for line in proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.sm.append(float(r[0])); self.mx = max(self.mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML indexes physical devices; honour CUDA_VISIBLE_DEVICES if it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.gpu
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[self.gpu])
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.source = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            pass
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._smi_loop, args=(self.proc,), daemon=True)
            self.thread.start()
        except Exception:
            self.source = None

    def stop(self):
        self.stop_flag.set()
        if getattr(self, "proc", None):
            self.proc.terminate()
        if self.thread:
            self.thread.join(timeout=1.0)
        sm = list(self.sm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.mx or None, "reasons": sorted(self.reasons), "samples": len(sm),
                "source": self.source}


def reference_throughput(n_timed: int, n_warm: int, workers: int | None = None, threads: int = 4):
    """Run the reference CPU path: `workers` processes x `threads` ggml threads, one prompt each, (n_warm + n_timed) repetitions.
    Default split = the best of the (workers, threads) grid measured on the GPU box (2 x Xeon 8562Y+, 128 logical CPUs):
    1x8 2.1 | 1x32 1.5 | 4x8 3.5 | 8x8 3.1 | 16x8 1.8 | 16x4 4.0 | 32x4 2.6 | 8x16 0.9 audio-s/s -> ncpu/8 workers x 4 threads
    (one ggml thread per physical core; its spin-wait barriers and memory traffic make wider splits slower)."""
    binp, build = ref_bin("kokoro_ref")
    if not os.path.exists(binp):
        return None, "oracle/_ref/kokoro_ref missing (run `make -C oracle ref` where /root/reference exists)"
    hc = host_cpus()
    ncpu = hc["usable"]
    threads = max(1, min(threads, ncpu))
    workers = workers or max(1, ncpu // (2 * threads))
    gguf = _gguf()
    prompts = _prompts(0)
    tmp = tempfile.mkdtemp(prefix="b2ref_")
    procs = []
    for w in range(workers):
        tok = os.path.join(tmp, f"tok{w}.txt")
        with open(tok, "w") as f:
            f.write(" ".join(map(str, prompts[w % len(prompts)])) + "\n")
        cmd = [binp, gguf, tok, os.path.join(tmp, f"o{w}"), "--threads", str(threads), "--reps", str(n_warm + n_timed), "--warm", str(n_warm), "--quiet"]
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    audio, wall = 0.0, 0.0
    for p in procs:
        out, err = p.communicate()
        for line in out.splitlines():
            if line.startswith("SUMMARY"):
                s = json.loads(line[len("SUMMARY "):])
                audio += s["audio_s"]; wall = max(wall, s["wall_s"])
    if wall <= 0:
        return None, "reference run produced no SUMMARY"
    return {"value": audio / wall, "unit": "audio-s/s", "cores": workers * threads, "kind": "reference", "wall_s": wall, "timed_runs_per_worker": n_timed,
            "workers": workers, "threads_per_worker": threads, "host": hc, "build": build,
            "sample": f"{workers} worker processes x {threads} ggml threads (of {hc['usable']} usable logical CPUs: affinity {hc['affinity']}, cgroup quota {hc['cgroup_quota_cpus']}, {hc['model']}), "
                      f"each {n_timed} timed run(s) of one 66-token prompt (~4.95 s audio, one of the batch's 32 prompts; the reference has no batching) after {n_warm} warm-up; "
                      f"throughput = total audio / slowest worker's wall; {build}"}, None


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # a step of this arm = every worker synthesising one prompt (a bounded sample of the batch-32 step: 16 of its prompts on this
    # box); capped at 6 timed repetitions so that any --steps finishes within a few minutes
    # the worker x thread split is swept on THIS box first (one timed run each, ~10 s per point), then the best split is timed
    ncpu = host_cpus()["usable"]
    grid = sorted({(max(1, ncpu // (2 * t)), t) for t in (2, 4, 8)} | {(max(1, ncpu // 4), 4)})
    sweep, best = [], None
    for w, t in grid:
        r, _ = reference_throughput(n_timed=1, n_warm=1, workers=w, threads=t)
        if r:
            sweep.append({"workers": w, "threads": t, "audio_s_per_s": round(r["value"], 3)})
            if best is None or r["value"] > best[0]:
                best = (r["value"], w, t)
    base, why = reference_throughput(n_timed=max(1, min(args.steps, 6)), n_warm=max(1, min(args.warmup, 1)), workers=best[1] if best else None, threads=best[2] if best else 4)
    if base is not None:
        base["split_sweep"] = sweep
    if base is None:
        print(json.dumps({"impl": "reference", "unavailable": why}))
        return 0
    line = {
        "impl": "reference", "metric": "audio_seconds_per_second", "value": base["value"], "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": base["timed_runs_per_worker"], "warmup": max(1, min(args.warmup, 1)), "ms_per_step": base["wall_s"] * 1e3 / base["timed_runs_per_worker"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 weights / f32 activations (GGML CPU)", "data": "synthetic",
        "config": {"workload": "Kokoro-82M fp16 GGUF (synthetic weights), 64-char (66-token) prompts, reference CPU GGML path, sequential per worker",
                   "sample_note": "a step of this arm = every worker synthesising ONE prompt of the batch-32 workload (the reference has no batching); at most 6 timed repetitions; same GGUF and prompts as the CUDA arm"},
        "cpu_baseline": base, "e2e": {"value": base["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def _dac_reference_sample(threads: int = 4):
    """-> (audio_s, wall_s, workers) of oracle/_ref/dac_ref decoding one 861-frame utterance per worker process, or None"""
    ref, _build = ref_bin("dac_ref")
    if not os.path.exists(ref):
        return None
    from tts_cpp_b200.synth import cached_dac_gguf, synthetic_codes
    gguf = cached_dac_gguf(seed=0, max_frames=870)
    ncpu = host_cpus()["usable"]
    workers = max(1, min(16, ncpu // (2 * threads)))
    tmp = tempfile.mkdtemp(prefix="b2dac_")
    procs = []
    for w, c in enumerate(synthetic_codes(workers, 861)):
        cf = os.path.join(tmp, f"c{w}.txt")
        open(cf, "w").write(" ".join(map(str, c.reshape(-1))) + "\n")
        procs.append(subprocess.Popen([ref, gguf, cf, os.path.join(tmp, f"o{w}"), "--threads", str(threads), "--quiet"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    audio, wall = 0.0, 0.0
    for p in procs:
        out, _ = p.communicate()
        for line in out.splitlines():
            if line.startswith("SUMMARY"):
                sm = json.loads(line[len("SUMMARY "):])
                audio += sm["audio_s"]; wall = max(wall, sm["wall_s"])
    return (audio, wall, workers) if wall > 0 else None


def run_dac_reference(args, threads: int = 4):
    """The reference's dac_runner (oracle/_ref/dac_ref) on the host cores: W worker processes x `threads` ggml threads, each decoding one
    861-frame utterance (10 s of audio) of the same synthetic GGUF -- a bounded sample of the batch-16 step."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    r = _dac_reference_sample(threads)
    if r is None:
        print(json.dumps({"impl": "reference", "workload": "dac", "unavailable": "oracle/_ref/dac_ref missing or silent (run `make -C oracle ref` where /root/reference exists)"}))
        return 0
    audio, wall, workers = r
    print(json.dumps({"impl": "reference", "metric": "audio_seconds_per_second", "workload": "DAC codec decode, 861-frame utterances (10 s @ 44.1 kHz), reference CPU GGML path",
                      "value": audio / wall, "unit": "audio-s/s", "n_gpus": args.gpus, "steps": 1, "ms_per_step": wall * 1e3,
                      "cpu_baseline": {"value": audio / wall, "unit": "audio-s/s", "cores": workers * threads, "kind": "reference",
                                       "sample": f"{workers} worker processes x {threads} ggml threads, one 10 s utterance each; throughput = total audio / slowest worker"},
                      "e2e": {"value": audio / wall, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "data": "synthetic"}))
    return 0


def run_parler_reference(args, threads: int = 4, sample_steps: int = 60):
    """The reference's Parler decode loop (oracle/_ref/parler_ref) + DAC decode (dac_ref) on the host cores, a bounded sample of config 3: every worker
    process generates `sample_steps` frames of one utterance; the DAC rate comes from the 10 s-utterance sample of run_dac_reference.  The two stages run one
    after the other in the reference, so the pipeline rate is the harmonic combination of the stage rates."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    ref, _build = ref_bin("parler_ref")
    if not os.path.exists(ref):
        print(json.dumps({"impl": "reference", "workload": "parler", "unavailable": "oracle/_ref/parler_ref missing (run `make -C oracle ref` where /root/reference exists)"}))
        return 0
    import numpy as np
    from tts_cpp_b200.synth import PARLER_MINI_SHAPE, cached_parler_gguf
    quant = None if args.parler_dtype == "f16" else args.parler_dtype.upper()
    gguf = cached_parler_gguf(seed=0, f16=quant is None, quant=quant, **PARLER_MINI_SHAPE)
    ncpu = host_cpus()["usable"]
    workers = max(1, min(16, ncpu // (2 * threads)))
    tmp = tempfile.mkdtemp(prefix="b2par_")
    rng = np.random.default_rng(5)
    procs = []
    for w in range(workers):
        pf = os.path.join(tmp, f"p{w}.txt")
        open(pf, "w").write(" ".join(map(str, rng.integers(1, 500, size=24))) + "\n")
        procs.append(subprocess.Popen([ref, gguf, pf, os.path.join(tmp, f"o{w}"), "--steps", str(sample_steps), "--threads", str(threads), "--quiet"],
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    frames, wall = 0, 0.0
    for p in procs:
        out, _ = p.communicate()
        for line in out.splitlines():
            if line.startswith("SUMMARY"):
                sm = json.loads(line[len("SUMMARY "):])
                frames += sm["steps"]; wall = max(wall, sm["wall_s"])
    dac = _dac_reference_sample(threads)
    if wall <= 0 or dac is None:
        print(json.dumps({"impl": "reference", "workload": "parler", "unavailable": "parler_ref / dac_ref produced no SUMMARY"}))
        return 0
    ar_rate = (frames * 512 / 44100.0) / wall                     # audio-seconds of frames generated per second (the prompt pass is inside wall)
    dac_rate = dac[0] / dac[1]
    rate = 1.0 / (1.0 / ar_rate + 1.0 / dac_rate)
    print(json.dumps({"impl": "reference", "metric": "audio_seconds_per_second", "workload": f"Parler-TTS-Mini-sized {args.parler_dtype} decoder (synthetic) greedy AR decode + DAC decode, reference CPU GGML path",
                      "value": rate, "unit": "audio-s/s", "n_gpus": args.gpus, "steps": 1, "ms_per_step": wall * 1e3, "ar_audio_s_per_s": ar_rate, "dac_audio_s_per_s": dac_rate,
                      "cpu_baseline": {"value": rate, "unit": "audio-s/s", "cores": workers * threads, "kind": "reference",
                                       "sample": f"{workers} worker processes x {threads} ggml threads: {sample_steps} decode steps of one utterance each (KV cache shorter than at 861 steps), "
                                                 "then one 10 s DAC decode each; stage rates combined harmonically"},
                      "e2e": {"value": rate, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "data": "synthetic"}))
    return 0


def run_dac(args):
    """Secondary line (not the headline): DAC codec decode, batch 16 x 861 frames (10 s @ 44.1 kHz each), synthetic F32 DAC GGUF."""
    if args.impl == "reference":
        return run_dac_reference(args)
    import torch  # noqa: F401  (device context / first-import cost, like the main arm)
    from tts_cpp_b200.binding import Context, dac_runner_from_file
    from tts_cpp_b200.synth import cached_dac_gguf, synthetic_codes
    ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    dac = dac_runner_from_file(cached_dac_gguf(seed=0, max_frames=64), ctx=ctx)
    B, frames = 16, 861
    codes = synthetic_codes(B, frames)
    for _ in range(max(args.warmup, 2)):
        pcm = dac.run_batch(codes, copy=False)
    audio_s = sum(p.shape[0] for p in pcm) / 44100.0
    l0 = ctx.launches()
    dev_ms, t0 = 0.0, time.perf_counter()
    for _ in range(args.steps):
        dac.run_batch(codes, copy=False)
        dev_ms += dac.last_ms()
    wall = time.perf_counter() - t0
    print(json.dumps({
        "metric": "audio_seconds_per_second", "workload": "DAC codec decode (SURVEY 8a-C), batch 16 x 861 frames (10 s @ 44.1 kHz), synthetic F32 DAC GGUF",
        "value": audio_s * args.steps / (dev_ms * 1e-3), "unit": "audio-s/s", "n_gpus": 1, "steps": args.steps, "ms_per_step": dev_ms / args.steps,
        "e2e": {"value": audio_s * args.steps / wall, "unit": "audio-s/s", "ms_per_step": wall * 1e3 / args.steps},
        "gpu_launches": int(ctx.launches() - l0), "gemm_dispatch": dict(zip(("tcgen05_tma", "mma_sync_fallback"), ctx.gemm_launches())),
        "dtype": "split-fp16 operands (fp32-faithful), f32 accumulate / activations", "data": "synthetic"}))
    return 0


def run_snac_reference(args, threads: int = 4):
    """The reference's snac_runner (oracle/_ref/snac_ref) on the host cores: W worker processes x `threads` ggml threads, each decoding one 468-fine-frame
    utterance (9.98 s of audio) of the same synthetic GGUF -- a bounded sample of the batch-16 step."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    ref, _build = ref_bin("snac_ref")
    if not os.path.exists(ref):
        print(json.dumps({"impl": "reference", "workload": "snac", "unavailable": "oracle/_ref/snac_ref missing (run `make -C oracle ref` where /root/reference exists)"}))
        return 0
    import numpy as np
    from tts_cpp_b200.synth import cached_snac_gguf, synthetic_snac_codes
    gguf = cached_snac_gguf(seed=0, max_frames=480)
    ncpu = host_cpus()["usable"]
    workers = max(1, min(16, ncpu // (2 * threads)))
    tmp = tempfile.mkdtemp(prefix="b2snac_")
    procs = []
    for w, c in enumerate(synthetic_snac_codes(workers, 468)):
        cf = os.path.join(tmp, f"c{w}.txt")
        open(cf, "w").write(" ".join(map(str, np.concatenate(c))) + "\n")
        procs.append(subprocess.Popen([ref, gguf, cf, os.path.join(tmp, f"o{w}"), "--threads", str(threads), "--quiet"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    audio, wall = 0.0, 0.0
    for p in procs:
        out, _ = p.communicate()
        for line in out.splitlines():
            if line.startswith("SUMMARY"):
                sm = json.loads(line[len("SUMMARY "):])
                audio += sm["audio_s"]; wall = max(wall, sm["wall_s"])
    if wall <= 0:
        print(json.dumps({"impl": "reference", "workload": "snac", "unavailable": "snac_ref produced no SUMMARY"}))
        return 0
    print(json.dumps({"impl": "reference", "metric": "audio_seconds_per_second", "workload": "SNAC codec decode, 468-fine-frame utterances (9.98 s @ 24 kHz), reference CPU GGML path",
                      "value": audio / wall, "unit": "audio-s/s", "n_gpus": args.gpus, "steps": 1, "ms_per_step": wall * 1e3,
                      "cpu_baseline": {"value": audio / wall, "unit": "audio-s/s", "cores": workers * threads, "kind": "reference",
                                       "sample": f"{workers} worker processes x {threads} ggml threads, one 9.98 s utterance each; throughput = total audio / slowest worker"},
                      "e2e": {"value": audio / wall, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "data": "synthetic"}))
    return 0


def run_snac(args):
    """Secondary line (not the headline): SNAC codec decode (Orpheus' codec), batch 16 x 468 fine frames (9.98 s @ 24 kHz each), synthetic F32 SNAC
    GGUF.  The reference's process-wide normal-noise stream is part of the call (generated on the host inside decode_batch, like snac_runner::set_inputs)."""
    if args.impl == "reference":
        return run_snac_reference(args)
    import torch  # noqa: F401  (device context / first-import cost, like the main arm)
    from tts_cpp_b200.binding import Context, snac_runner_from_file
    from tts_cpp_b200.synth import cached_snac_gguf, synthetic_snac_codes
    ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    snac = snac_runner_from_file(cached_snac_gguf(seed=0, max_frames=480), ctx=ctx)
    B, fine = 16, 468
    codes = synthetic_snac_codes(B, fine, codebook=snac.codebook_size)
    for _ in range(max(args.warmup, 2)):
        pcm = snac.run_batch(codes, copy=False)
    audio_s = sum(p.shape[0] for p in pcm) / 24000.0
    l0 = ctx.launches()
    dev_ms, t0 = 0.0, time.perf_counter()
    for _ in range(args.steps):
        snac.run_batch(codes, copy=False)
        dev_ms += snac.last_ms()
    wall = time.perf_counter() - t0
    print(json.dumps({
        "metric": "audio_seconds_per_second", "workload": "SNAC codec decode (SURVEY 8a-C), batch 16 x 468 fine frames (9.98 s @ 24 kHz), synthetic F32 SNAC GGUF",
        "value": audio_s * args.steps / (dev_ms * 1e-3), "unit": "audio-s/s", "n_gpus": 1, "steps": args.steps, "ms_per_step": dev_ms / args.steps,
        "e2e": {"value": audio_s * args.steps / wall, "unit": "audio-s/s", "ms_per_step": wall * 1e3 / args.steps,
                "note": "includes the host-side libstdc++-compatible normal noise stream (840 floats per fine frame)"},
        "gpu_launches": int(ctx.launches() - l0), "gemm_dispatch": dict(zip(("tcgen05_tma", "mma_sync_fallback"), ctx.gemm_launches())),
        "dtype": "split-fp16 operands (fp32-faithful), f32 accumulate / activations", "data": "synthetic"}))
    return 0


def run_parler(args):
    """Secondary line (not the headline; written before it could be run on a B200): BASELINE config 3's shape -- a Parler-TTS-Mini-sized F16 decoder
    (synthetic weights), batch 16, 10 s of audio per utterance (861 DAC frames + the 8-step delay tail), greedy, then the DAC decode of the frames.
    Random weights emit special ids at random, which the reference would drop frame by frame: they are folded into the codebook range here so that every
    utterance decodes its full 10 s (said in `config`)."""
    if args.impl == "reference":
        return run_parler_reference(args)
    import numpy as np
    import torch  # noqa: F401  (device context / first-import cost, like the main arm)
    from tts_cpp_b200.ar_host import parler_adjust_output_tokens
    from tts_cpp_b200.binding import Context, dac_runner_from_file, parler_runner_from_file
    from tts_cpp_b200.synth import PARLER_MINI_SHAPE, cached_dac_gguf, cached_parler_gguf
    ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    quant = None if args.parler_dtype == "f16" else args.parler_dtype.upper()
    par = parler_runner_from_file(cached_parler_gguf(seed=0, f16=quant is None, quant=quant, **PARLER_MINI_SHAPE), ctx=ctx)
    dac = dac_runner_from_file(cached_dac_gguf(seed=0, max_frames=64), ctx=ctx)
    B, frames = 16, 861
    n_steps = frames + par.n_heads - 1
    rng = np.random.default_rng(5)
    prompts = [rng.integers(1, 500, size=24).astype(np.uint32) for _ in range(B)]

    def step():
        toks = par.generate_greedy(prompts, n_steps)
        t_ar = par.last_ms()
        codes = [parler_adjust_output_tokens(t % 1024, 1024) for t in toks]
        pcm = dac.run_batch(codes, copy=False)
        return t_ar, dac.last_ms(), sum(p.shape[0] for p in pcm) / 44100.0

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    l0 = ctx.launches()
    ar_ms = dac_ms = audio_s = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        a, d, au = step()
        ar_ms += a; dac_ms += d; audio_s += au
    wall = time.perf_counter() - t0
    dev_ms = ar_ms + dac_ms
    _, hbm, peak_src = _peaks()
    per_step_ms = ar_ms / args.steps / n_steps
    # decode-step roofline, SURVEY 8(d): W_step (every weight tensor a step touches, once, stored dtype) + per sequence the self-attention KV read up to the current
    # position and the new row written, at 2 bytes per element (the metric's definition, independent of how the cache is stored), + the logits; averaged over the steps
    L, H = par.n_layers, par.hidden_size
    w_step = par.step_weight_bytes()
    p_avg = 24 + (n_steps - 1) / 2.0
    kv_read, kv_write, lg = 2 * L * p_avg * H * 2, 2 * L * H * 2, par.n_heads * par.out_vocab * 4
    alg = w_step + B * (kv_read + kv_write + lg)
    pk_launches, pk_steps = par.pdk_stats()
    print(json.dumps({
        "metric": "audio_seconds_per_second", "workload": f"Parler-TTS-Mini-sized {args.parler_dtype} decoder (synthetic), batch 16 x 10 s, greedy AR decode + DAC decode (BASELINE config 3)",
        "value": audio_s / (dev_ms * 1e-3), "unit": "audio-s/s", "n_gpus": 1, "steps": args.steps, "ms_per_step": dev_ms / args.steps,
        "ar_ms_per_step": ar_ms / args.steps, "dac_ms_per_step": dac_ms / args.steps, "decode_step_ms": per_step_ms,
        "e2e": {"value": audio_s / wall, "unit": "audio-s/s", "ms_per_step": wall * 1e3 / args.steps},
        "roofline": {"bound": "hbm", "kernel": "pdk_kernel (persistent decode kernel)" if pk_steps else "launch-per-op decode step", "achieved": alg / (per_step_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": alg / (per_step_ms * 1e-3) / 1e9 / hbm, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_decode_step": alg, "terms": {"W_step": w_step, "kv_read_per_seq_avg": kv_read, "kv_write_per_seq": kv_write, "logits_per_seq": lg, "batch": B, "avg_position": p_avg},
                     "weights_only_frac": w_step / (per_step_ms * 1e-3) / 1e9 / hbm,
                     "note": "decode step incl. the prompt pass amortised over the 869 steps; KV term at 2 B / element per SURVEY 8(d)"},
        "gpu_launches": int(ctx.launches() - l0), "persistent_kernel": {"launches": pk_launches, "decode_steps": pk_steps},
        "dtype": ("f16 matrices x fp16-rounded activations, f32 accumulate" if quant is None else f"{quant} blocks x Q8_0-requantised activations, int32 block dots, f32 accumulate") + " (the reference's numerics for this GGUF)",
        "data": "synthetic", "config": {"workload": f"parler-mini {args.parler_dtype} (24 layers x 1024, 9 codebooks), batch 16, 869 decode steps -> 861 frames, special ids folded mod 1024, DAC 44.1 kHz decode",
                                         "switches": {k: os.environ.get(k, d) for k, d in (("B2TTS_AR_PDK", "1"), ("B2TTS_KV", "f16"), ("B2TTS_AR_FUSE", "1"), ("B2TTS_AR_GRAPH", "1"), ("B2TTS_AR_MMA", "1"), ("B2TTS_AR_ATT", "gqa"))}}}))
    return 0


def _sharded(args):
    """-> (rank, world, dist or None): configs 4 / 5 shard independent utterances over the GPUs of a box: one process per GPU under torchrun, every rank runs
    its own per-GPU batch, no data-path collective; NCCL carries the barriers and the max-over-ranks of the device-timed milliseconds."""
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        return 0, 1, None
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, dist


def _max_over_ranks(dist, ms: float) -> float:
    if dist is None:
        return ms
    import torch
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_orpheus_sharded(args):
    """BASELINE config 5 as it is stated -- Orpheus-3B q8_0, 64 utterances sharded over 8 GPUs = 8 per GPU (N GPUs: 8 N utterances, weak scaling): every rank decodes its 8
    sequences for 168 steps (24 SNAC frames each) inside the persistent kernel; barrier + synchronize on both sides, device-timed, max over ranks."""
    import torch
    from tts_cpp_b200.binding import Context
    from tts_cpp_b200.synth import build_orpheus_direct
    rank, world, dist = _sharded(args)
    ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    orph = build_orpheus_direct(ctx, dtype=args.orpheus_dtype)
    B, n_tokens = 8, 7 * 24
    rng = np.random.default_rng(9 + rank)
    prompts = [rng.integers(1, 100000, size=40).astype(np.uint32) for _ in range(B)]
    orph.generate_greedy(prompts, 16)
    ms = []
    for _ in range(max(1, min(args.steps, 3))):
        if dist is not None: dist.barrier()
        torch.cuda.synchronize()
        orph.generate_greedy(prompts, n_tokens)
        torch.cuda.synchronize()
        ms.append(_max_over_ranks(dist, orph.last_ms()))
    if rank == 0:
        best = min(ms)
        audio = world * B * (n_tokens / 7) * (2048 / 24000.0)
        print(json.dumps({"metric": "audio_seconds_per_second", "workload": f"Orpheus-3B-shaped {args.orpheus_dtype} decoder (synthetic), greedy AR decode, {world * B} utterances sharded over {world} GPU(s), 8 per GPU (BASELINE config 5)",
                          "value": audio / (best * 1e-3), "unit": "audio-s/s", "n_gpus": world, "steps": len(ms), "ms_per_step": best, "ms_per_decode_step": best / n_tokens, "scaling": "weak", "collectives": None,
                          "persistent_kernel": dict(zip(("launches", "steps"), orph.pdk_stats())), "data": "synthetic", "config": {"workload": "orpheus-3b shape, 40-token prompts, 168 decode steps, 8 sequences per GPU"}}))
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    return 0


def run_orpheus(args):
    """Secondary line: BASELINE config 5's model on ONE GPU of the 8 -- an Orpheus-3B-shaped decoder (28 layers x 3072, 24 / 8 heads x 128, ffn 8192, vocab 156 940) with
    Q8_0 matrices (our own writer: the reference's quantize tool refuses Orpheus and its runtime is F32-only), random weights handed over tensor by tensor (no GGUF file),
    greedy, decode steps inside the persistent decode kernel (pdk.cuh: int8 MMA over activations quantised per 32-block, ggml_vec_dot_q8_0_q8_0's arithmetic;
    B2TTS_AR_PDK=0: launch-per-op dp4a path under CUDA-graph replay).  Sweep of the per-GPU batch {1, 2, 4, 8, 16} (config 5: 64 utterances over 8 GPUs = 8 per GPU);
    a "step" = `n_tokens` decode steps of the whole batch (7 tokens = one 85.3 ms SNAC frame).  --orpheus-dtype f16: the same shape with F16 matrices, whose decode
    steps run inside the persistent decode kernel (pdk.cuh)."""
    if args.impl == "reference":
        print(json.dumps({"impl": "reference", "workload": "orpheus", "unavailable": "the reference runs Orpheus in F32 only (README.md:25): 13 GB of weights per CPU worker, minutes per second of audio; not timed here"}))
        return 0
    import torch  # noqa: F401
    from tts_cpp_b200.binding import Context
    from tts_cpp_b200.synth import build_orpheus_direct
    ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    t0 = time.perf_counter()
    orph = build_orpheus_direct(ctx, dtype=args.orpheus_dtype)
    load_s = time.perf_counter() - t0
    n_tokens = 7 * 24                                           # 24 SNAC frames = 2.05 s of audio per sequence per step
    rng = np.random.default_rng(9)
    _, hbm, peak_src = _peaks()
    w_step = orph.step_weight_bytes()
    L, KV = orph.n_layers, 1024
    sweep = []
    for B in (1, 2, 4, 8, 16):
        prompts = [rng.integers(1, 100000, size=40).astype(np.uint32) for _ in range(B)]
        orph.generate_greedy(prompts, 16)                       # warm-up (graph instantiation, arena)
        ms = []
        for _ in range(max(1, min(args.steps, 3))):
            orph.generate_greedy(prompts, n_tokens)
            ms.append(orph.last_ms())
        step_ms = min(ms) / n_tokens
        pos = 40 + n_tokens / 2.0
        alg = w_step + B * (2 * L * pos * KV * 2 + 2 * L * KV * 2 + orph.vocab_size * 4)
        sweep.append({"batch_per_gpu": B, "ms_per_decode_step": step_ms, "audio_s_per_s": B * (n_tokens / 7) * (2048 / 24000.0) / (min(ms) * 1e-3),
                      "algorithmic_bytes_per_step": alg, "achieved_gbs": alg / (step_ms * 1e-3) / 1e9, "frac": alg / (step_ms * 1e-3) / 1e9 / hbm})
    best = next(x for x in sweep if x["batch_per_gpu"] == 8)
    print(json.dumps({"metric": "audio_seconds_per_second", "workload": f"Orpheus-3B-shaped {args.orpheus_dtype} decoder (synthetic), greedy AR decode, per-GPU batch sweep (BASELINE config 5: 8 per GPU x 8 GPUs); SNAC decode measured by --workload snac",
                      "value": best["audio_s_per_s"], "unit": "audio-s/s", "n_gpus": 1, "steps": args.steps, "ms_per_step": best["ms_per_decode_step"] * n_tokens,
                      "persistent_kernel": dict(zip(("launches", "steps"), orph.pdk_stats())),
                      "sweep": sweep, "roofline": {"bound": "hbm", "kernel": "pdk_kernel (persistent decode kernel)" if orph.pdk_stats()[1] else "launch-per-op decode step (gemv_rows_q_kernel dp4a / attention_gqa_kernel)", "achieved": best["achieved_gbs"], "peak": hbm, "unit": "GB/s",
                                                   "frac": best["frac"], "traffic": None, "peak_source": peak_src, "W_step": w_step},
                      "load_s": load_s, "weight_bytes": orph.weight_bytes(), "dtype": f"{args.orpheus_dtype} matrices (Q8_0: Q8_0-requantised activations, int32 block dots, f32 accumulate)", "data": "synthetic",
                      "config": {"workload": "orpheus-3b shape, 40-token prompts, 168 decode steps per timed generation, greedy"}}))
    return 0


def run_dia_sharded(args):
    """BASELINE config 4 as it is stated -- Dia-1.6B F16, 8 utterances sharded over 4 GPUs = 2 per GPU (N GPUs: 2 N utterances, weak scaling): every rank generates 10 s for
    its two CFG pairs inside the persistent kernel and decodes the frames with its DAC; device-timed per rank, max over ranks, barriers on both sides."""
    import torch
    from tts_cpp_b200.ar_host import dia_adjust_output_tokens
    from tts_cpp_b200.binding import Context, dac_runner_from_file
    from tts_cpp_b200.synth import build_dia_direct, cached_dac_gguf
    rank, world, dist = _sharded(args)
    ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    dia = build_dia_direct(ctx, dtype="f16")
    if rank == 0: cached_dac_gguf(seed=0, max_frames=64)
    if dist is not None: dist.barrier()
    dac = dac_runner_from_file(cached_dac_gguf(seed=0, max_frames=64), ctx=ctx)
    B, frames = 2, 861
    n_steps = frames + 15
    rng = np.random.default_rng(4 + rank)
    prompts = [np.concatenate([[1], rng.integers(32, 127, size=62), [2], rng.integers(32, 127, size=64)]).astype(np.uint32) for _ in range(B)]

    def step():
        toks, ngen = dia.generate_greedy(prompts, n_steps)
        t_ar = dia.last_ms()
        codes = [dia_adjust_output_tokens(t % 1024, 1024) for t in toks]
        pcm = dac.run_batch(codes, copy=False)
        return t_ar + dac.last_ms(), sum(p.shape[0] for p in pcm) / 44100.0

    step()
    ms, audio = [], 0.0
    for _ in range(max(1, min(args.steps, 3))):
        if dist is not None: dist.barrier()
        torch.cuda.synchronize()
        t, audio = step()
        ms.append(_max_over_ranks(dist, t))
    if rank == 0:
        best = min(ms)
        print(json.dumps({"metric": "audio_seconds_per_second", "workload": f"Dia-1.6B-shaped F16 model (synthetic), {world * B} utterances (CFG pairs) sharded over {world} GPU(s), 2 per GPU, 10 s each, greedy AR decode + DAC decode (BASELINE config 4)",
                          "value": world * audio / (best * 1e-3), "unit": "audio-s/s", "n_gpus": world, "steps": len(ms), "ms_per_step": best, "scaling": "weak", "collectives": None,
                          "persistent_kernel": dict(zip(("launches", "steps"), dia.pdk_stats())), "data": "synthetic", "config": {"workload": "dia-1.6b shape, 128-byte prompts, 876 decode steps, 2 utterances per GPU"}}))
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    return 0


def run_dia(args):
    """Secondary line: BASELINE config 4's model on ONE GPU of the 4 -- a Dia-1.6B-shaped F16 model (encoder 12 x 1024, decoder 18 x 2048, 16 q / 4 kv heads x 128, ffn 8192),
    2 utterances per GPU (8 over 4 GPUs), each a CFG pair, 128-byte two-speaker prompts padded to the 1 024-position encoder context, 10 s of audio (861 frames + the
    15-step delay tail), greedy, the decoder loop inside the persistent decode kernel (B2TTS_AR_PDK=0: launch-per-op path), then the DAC decode of the frames.  Under torchrun:
    run_dia_sharded."""
    if args.impl == "reference":
        print(json.dumps({"impl": "reference", "workload": "dia", "unavailable": "not timed: a 1.6 B-parameter CPU decode of 876 steps per worker takes tens of minutes; the Parler arm (--workload parler --impl reference) is the timed AR reference"}))
        return 0
    import torch  # noqa: F401
    from tts_cpp_b200.ar_host import dia_adjust_output_tokens
    from tts_cpp_b200.binding import Context, dac_runner_from_file
    from tts_cpp_b200.synth import build_dia_direct, cached_dac_gguf
    ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    t0 = time.perf_counter()
    dia = build_dia_direct(ctx, dtype="f16")
    load_s = time.perf_counter() - t0
    dac = dac_runner_from_file(cached_dac_gguf(seed=0, max_frames=64), ctx=ctx)
    B, frames = 2, 861
    n_steps = frames + 15
    rng = np.random.default_rng(4)
    prompts = [np.concatenate([[1], rng.integers(32, 127, size=62), [2], rng.integers(32, 127, size=64)]).astype(np.uint32) for _ in range(B)]

    def step():
        toks, ngen = dia.generate_greedy(prompts, n_steps)
        t_ar = dia.last_ms()
        codes = [dia_adjust_output_tokens(t % 1024, 1024) for t in toks]
        pcm = dac.run_batch(codes, copy=False)
        return t_ar, dac.last_ms(), sum(p.shape[0] for p in pcm) / 44100.0, int(ngen.min())

    step()
    ar_ms = dac_ms = audio_s = 0.0
    t0 = time.perf_counter()
    for _ in range(max(1, min(args.steps, 3))):
        a, d, au, ng = step()
        ar_ms += a; dac_ms += d; audio_s += au
    n = max(1, min(args.steps, 3))
    wall = time.perf_counter() - t0
    _, hbm, peak_src = _peaks()
    step_ms = ar_ms / n / n_steps
    print(json.dumps({"metric": "audio_seconds_per_second", "workload": "Dia-1.6B-shaped F16 model (synthetic), 2 utterances (CFG pairs) per GPU, 10 s each, greedy AR decode + DAC decode (BASELINE config 4, one of its 4 GPUs)",
                      "value": audio_s / ((ar_ms + dac_ms) * 1e-3), "unit": "audio-s/s", "n_gpus": 1, "steps": n, "ms_per_step": (ar_ms + dac_ms) / n, "ar_ms_per_step": ar_ms / n, "dac_ms_per_step": dac_ms / n,
                      "decode_step_ms": step_ms, "frames_generated_min": ng, "e2e": {"value": audio_s / wall, "unit": "audio-s/s"},
                      "persistent_kernel": dict(zip(("launches", "steps"), dia.pdk_stats())),
                      "roofline": {"bound": "hbm", "kernel": "pdk_kernel (persistent decode kernel)" if dia.pdk_stats()[1] else "launch-per-op decode step (gemv_mma_kernel / attention_gqa_kernel)",
                                   "achieved": dia.weight_bytes() / (step_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                   "frac": dia.weight_bytes() / (step_ms * 1e-3) / 1e9 / hbm, "traffic": None, "peak_source": peak_src,
                                   "note": "algorithmic bytes = the resident weights once per step (encoder weights included: an upper bound of W_step by ~10 %); KV excluded"},
                      "load_s": load_s, "weight_bytes": dia.weight_bytes(), "dtype": "f16 matrices x fp16-rounded activations, f32 accumulate", "data": "synthetic",
                      "config": {"workload": "dia-1.6b shape, 128-byte prompts, 876 decode steps, EOS / PAD rows of the heads zeroed so that every generation runs its full length"}}))
    return 0


def decode_step_record(ctx):
    """The metric's second half, "decode-step HBM % peak": one decode step of BASELINE config 3's model (Parler-TTS-Mini-shaped F16 decoder, batch 16) around position 450,
    through the persistent decode kernel, against SURVEY 8(d)'s algorithmic bytes (W_step + per sequence the KV cache read up to the position and one row written, 2 B per
    element, + the logits).  Timed as the difference of two greedy generations (416 and 480 steps, CUDA events inside the library), i.e. over steps 416..479."""
    from tts_cpp_b200.binding import parler_runner_from_file
    from tts_cpp_b200.synth import PARLER_MINI_SHAPE, cached_parler_gguf
    par = parler_runner_from_file(cached_parler_gguf(seed=0, f16=True, **PARLER_MINI_SHAPE), ctx=ctx)
    B, n_prompt, n0, n1 = 16, 24, 416, 480
    rng = np.random.default_rng(5)
    prompts = [rng.integers(1, 500, size=n_prompt).astype(np.uint32) for _ in range(B)]
    par.generate_greedy(prompts, 64)                          # warm-up
    t = []
    for n in (n0, n1, n0, n1):
        par.generate_greedy(prompts, n)
        t.append(par.last_ms())
    ms = ((t[1] - t[0]) + (t[3] - t[2])) / 2.0 / (n1 - n0)
    L, H = par.n_layers, par.hidden_size
    pos = n_prompt + (n0 + n1 - 1) / 2.0
    w_step = par.step_weight_bytes()
    kv_read, kv_write, lg = 2 * L * pos * H * 2, 2 * L * H * 2, par.n_heads * par.out_vocab * 4
    alg = w_step + B * (kv_read + kv_write + lg)
    _, hbm, peak_src = _peaks()
    launches, steps = par.pdk_stats()
    rec = {"workload": "Parler-TTS-Mini-sized F16 decoder (synthetic), batch 16, greedy, decode steps 416..479 (positions ~440-504)", "ms_per_decode_step": ms,
           "algorithmic_bytes_per_step": alg, "terms": {"W_step": w_step, "kv_read_per_seq": kv_read, "kv_write_per_seq": kv_write, "logits_per_seq": lg, "batch": B, "position": pos},
           "achieved_gbs": alg / (ms * 1e-3) / 1e9, "peak_gbs": hbm, "frac": alg / (ms * 1e-3) / 1e9 / hbm, "peak_source": peak_src,
           "kernel": "pdk_kernel (persistent cooperative decode kernel: TMA weight ring, paged fp16 KV cache)" if steps else "launch-per-op decode path",
           "persistent_kernel": {"launches": launches, "decode_steps": steps}, "audio_s_per_s_ar_only": B * 512 / 44100.0 / (ms * 1e-3),
           "dram_traffic_per_step_ncu": _decode_traffic()}
    par.close()
    return rec


def strong_scaling_record(args, torch, dist, runner, rank, world, local):
    """SURVEY 8(d)/(e), the north star's multi-GPU sentence: a FIXED job of 32 utterances on rank 0's host -> NCCL scatter of the prompts -> every rank's batched forward
    (PCM stays in device memory) -> NCCL gather of the PCM to rank 0 -> one device-to-pinned-host copy there; everything inside the timed region, barrier + synchronize on
    both sides, max over ranks.  At N = 1 the same path without the collectives (the table's first row)."""
    from tts_cpp_b200 import sharding
    device = torch.device("cuda", local)
    counts = [BATCH // world + (1 if r < BATCH % world else 0) for r in range(world)]
    all_prompts = _prompts(0) if rank == 0 else None
    host_out = torch.empty(BATCH * 600 * 260, dtype=torch.float32).pin_memory() if rank == 0 else None
    split = [0.0, 0.0, 0.0]
    result = {}

    def step(timed):
        t0 = time.perf_counter()
        if dist is not None:
            mine = sharding.scatter_tokens_nccl(dist, torch, device, all_prompts, counts, src=0)
        else:
            mine = all_prompts
        t1 = time.perf_counter()
        ptr, stride, ns = runner.run_batch_device(mine) if mine else (0, 0, [])
        t2 = time.perf_counter()
        block = sharding.device_block(torch, device, ptr, len(ns), stride) if ns else None
        if dist is not None:
            got = sharding.gather_pcm_nccl(dist, torch, device, block, ns, counts, dst=0, host_out=host_out)
        else:
            packed = torch.cat([block[b, :n] for b, n in enumerate(ns)])
            host_out[:packed.numel()].copy_(packed, non_blocking=True)
            torch.cuda.synchronize()
            got = (host_out[:packed.numel()], ns)
        t3 = time.perf_counter()
        if timed:
            split[0] += t1 - t0; split[1] += t2 - t1; split[2] += t3 - t2
        if rank == 0:
            result["samples"] = int(sum(got[1])); result["utterances"] = len(got[1])
        return got

    for _ in range(2):
        step(False)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(True)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    wall = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([wall], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t[0])
    if rank != 0:
        return None
    audio_s = result["samples"] / 24000.0
    return {"scaling": "strong", "utterances_total": result["utterances"], "utterances_per_gpu": counts, "value": audio_s * args.steps / wall, "unit": "audio-s/s", "ms_per_step": wall * 1e3 / args.steps,
            "rank0_ms_per_step": {"scatter": split[0] * 1e3 / args.steps, "forward": split[1] * 1e3 / args.steps, "gather_and_d2h": split[2] * 1e3 / args.steps},
            "h2d_bytes_per_step": BATCH * (N_PHON + 2 + 1) * 8, "d2h_bytes_per_step": result["samples"] * 4,
            "collectives": None if dist is None else "NCCL: broadcast of the packed prompts (lengths + ids), all_gather of the PCM lengths, exact-length send/recv of the PCM from device memory to rank 0"}


def _decode_traffic():
    """DRAM bytes per decode step from the committed ncu capture of the persistent kernel (profiles/*_pdk_traffic.json), or None"""
    try:
        import glob
        cand = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_pdk_traffic.json")))
        if cand:
            tj = json.load(open(cand[-1]))
            return {"bytes": tj["dram_bytes_per_step"], "source": os.path.basename(cand[-1])}
    except Exception:
        pass
    return None


T5_SHAPE = dict(f16=True, layers=24, heads=16, ffn=2816, vocab=2048, out_size=1024, context_length=512)      # flan-t5-large's encoder + the projection to Parler-Mini's width


def _t5_prompts(B: int, n: int):
    import numpy as np
    rng = np.random.default_rng(0)
    return [[int(t) for t in rng.integers(2, 2048, n - 1)] + [1] for _ in range(B)]


def run_t5_reference(args, threads: int = 8, n_tokens: int = 16):
    """CPU arm of --workload t5: the unmodified t5_runner::run (oracle/_ref/t5_ref) on the same GGUF and prompt; per-pass time = (5 passes - 1 pass) / 4 of one process
    (removes the load).  A reported baseline."""
    from tts_cpp_b200.synth import cached_t5_gguf
    exe, build = ref_bin("t5_ref")
    if not os.path.exists(exe):
        print(json.dumps({"impl": "reference", "workload": "t5", "unavailable": "oracle/_ref/t5_ref not built"}))
        return 0
    g = cached_t5_gguf(**T5_SHAPE)
    p = ",".join(str(t) for t in _t5_prompts(1, n_tokens)[0])
    out = os.path.join(tempfile.gettempdir(), f"t5_ref_{os.getpid()}.bin")
    subprocess.run([exe, g, out, str(threads), p], check=True, stdout=subprocess.DEVNULL)          # warm the page cache
    t0 = time.perf_counter(); subprocess.run([exe, g, out, str(threads), p], check=True, stdout=subprocess.DEVNULL); t1 = time.perf_counter() - t0
    t0 = time.perf_counter(); subprocess.run([exe, g, out, str(threads)] + [p] * 5, check=True, stdout=subprocess.DEVNULL); t5 = time.perf_counter() - t0
    os.remove(out)
    ms = max((t5 - t1) / 4 * 1e3, 1e-3)
    cpu = {"value": 1e3 / ms, "unit": "encodings/s", "cores": threads, "kind": "reference", "build": build,
           "sample": f"one {n_tokens}-token prompt, 4 timed passes of t5_runner::run inside one process ({threads} ggml threads)"}
    print(json.dumps({"impl": "reference", "metric": "t5_prompt_encodings_per_second", "workload": f"T5 conditional-prompt encoder pass, flan-t5-large shape, F16 layer matrices, {n_tokens}-token prompt",
                      "value": 1e3 / ms, "unit": "encodings/s", "n_gpus": 1, "steps": 4, "ms_per_step": ms, "higher_is_better": True, "data": "synthetic", "cpu_baseline": cpu,
                      "e2e": {"value": 1e3 / ms, "unit": "encodings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
    return 0


def run_t5(args, n_tokens: int = 16):
    """Secondary line: one T5 conditional-prompt encoder pass (b2tts_t5_encode) at the flan-t5-large shape; value = passes per second of a 16-token voice
    description (device time, CUDA events around the forward); e2e = the same through the C-ABI with host token ids in and the host encoding out."""
    if args.impl == "reference":
        return run_t5_reference(args, n_tokens=n_tokens)
    import ctypes as C
    from tts_cpp_b200.binding import Context, lib, t5_runner_from_file
    from tts_cpp_b200.synth import cached_t5_gguf
    ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    t5 = t5_runner_from_file(cached_t5_gguf(**T5_SHAPE), ctx=ctx)
    lib().b2tts_t5_weight_bytes.restype = C.c_size_t
    wbytes = int(lib().b2tts_t5_weight_bytes(t5.h))
    prompts = _t5_prompts(1, n_tokens)
    for _ in range(max(args.warmup, 3)):
        t5.run(prompts)
    l0 = ctx.launches()
    dev_ms, t0 = 0.0, time.perf_counter()
    for _ in range(args.steps):
        t5.run(prompts)
        dev_ms += t5.last_ms()
    wall = time.perf_counter() - t0
    ms = dev_ms / args.steps
    act_bytes = n_tokens * (6 * 1024 + 2 * 2816 + 1024) * 4
    _, hbm, peak_src = _peaks()
    ach = (wbytes + act_bytes) / (ms * 1e-3) / 1e9
    print(json.dumps({
        "metric": "t5_prompt_encodings_per_second", "workload": f"T5 conditional-prompt encoder pass (SURVEY 8f row 3), flan-t5-large shape, F16 layer matrices, {n_tokens}-token prompt",
        "value": 1e3 / ms, "unit": "encodings/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "e2e": {"value": args.steps / wall, "unit": "encodings/s", "h2d_bytes_per_step": n_tokens * 4, "d2h_bytes_per_step": n_tokens * 1024 * 4},
        "gpu_launches": int(ctx.launches() - l0), "tensor_core_gemm": bool(lib().b2tts_t5_last_used_gemm(t5.h)),
        "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None,
                     "algorithmic_bytes": wbytes + act_bytes, "peak_source": peak_src, "note": "every weight once per pass + rows x (6 hidden + 2 ffn + out) x 4 B of activations"},
        "dtype": "f16 weights, fp16-rounded activations into exact products, f32 accumulate (the reference's F16 mul_mat numerics)", "data": "synthetic"}))
    t5.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling record (32 utterances in total: NCCL scatter of the prompts, forward, NCCL gather of the PCM to rank 0)")
    ap.add_argument("--no-decode-step", action="store_true", help="skip the decode-step HBM record of the default line (Parler-Mini F16, batch 16; ~30 s incl. writing its synthetic GGUF)")
    ap.add_argument("--parler-dtype", default="f16", choices=["f16", "q8_0", "q5_0", "q4_0"],
                    help="--workload parler: dtype of the decoder matrices (f16 = BASELINE config 3; q5_0 is what the reference's published Parler numbers use)")
    ap.add_argument("--orpheus-dtype", default="q8_0", choices=["q8_0", "f16", "f32"], help="--workload orpheus: dtype of the matrices (q8_0 = BASELINE config 5)")
    ap.add_argument("--workload", default="kokoro", choices=["kokoro", "dac", "snac", "parler", "orpheus", "dia", "t5"],
                    help="kokoro (default, the headline metric) | dac: codec decode of BASELINE config 3's shape (batch 16 x 10 s), a secondary line | "
                         "parler: config 3 end to end (AR decode + DAC), plain first path")
    args = ap.parse_args()
    if args.workload == "orpheus" and int(os.environ.get("WORLD_SIZE", "1")) > 1 and args.impl != "reference":
        return run_orpheus_sharded(args)
    if args.workload == "dia" and int(os.environ.get("WORLD_SIZE", "1")) > 1 and args.impl != "reference":
        return run_dia_sharded(args)
    if args.workload != "kokoro" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        if int(os.environ.get("RANK", "0")) == 0:       # the other secondary lines are single-GPU measurements; the headline workload is the one that shards
            print(json.dumps({"workload": args.workload, "unavailable": "secondary workloads are measured on one GPU (run without torchrun)"}))
        return 0
    if args.workload == "snac":
        return run_snac(args)
    if args.workload == "dac":
        return run_dac(args)
    if args.workload == "parler":
        return run_parler(args)
    if args.workload == "orpheus":
        return run_orpheus(args)
    if args.workload == "dia":
        return run_dia(args)
    if args.workload == "t5":
        return run_t5(args)
    if args.impl == "reference":
        return run_reference_arm(args)

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from tts_cpp_b200.binding import Context, runner_from_file, lib
    import ctypes as C

    gguf = _gguf() if rank == 0 else None
    if dist is not None:
        dist.barrier()
        gguf = _gguf()
    ctx = Context(local)
    runner = runner_from_file(gguf, ctx=ctx)
    prompts = _prompts(rank)
    h2d_bytes = sum(len(p) for p in prompts) * 4

    def step():
        pcms, durs = runner.run_batch(prompts, copy=False)   # PCM stays in the runner's pinned host buffers
        return pcms

    for _ in range(max(args.warmup, 3)):
        pcms = step()
    audio_s = sum(p.shape[0] for p in pcms) / 24000.0
    d2h_bytes = sum(p.shape[0] for p in pcms) * 4

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib().b2tts_prof_enable(ctx.h, 1)
    l0 = ctx.launches()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    dev_ms = 0.0
    pass_ms = [0.0, 0.0]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
        tm = runner.timings()
        dev_ms += tm["duration_ms"] + tm["generation_ms"]
        pass_ms[0] += tm["duration_ms"]; pass_ms[1] += tm["generation_ms"]
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    wall = time.perf_counter() - t0
    launches = ctx.launches() - l0
    clocks = sampler.stop() if rank == 0 else None

    prof = {}
    for kind, name in ((0, "conv_gemm"), (1, "bilstm"), (2, "norm_adain"), (3, "conv_transpose")):
        ms, fl, by, n = C.c_double(), C.c_double(), C.c_double(), C.c_uint64()
        lib().b2tts_prof_read(ctx.h, kind, C.byref(ms), C.byref(fl), C.byref(by), C.byref(n))
        prof[name] = {"ms": ms.value, "flops": fl.value, "bytes": by.value, "launches": n.value}
    lib().b2tts_prof_enable(ctx.h, 0)

    if dist is not None:
        t = torch.tensor([wall, dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall, dev_ms = float(t[0]), float(t[1])
        a = torch.tensor([audio_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(a, op=dist.ReduceOp.SUM)
        audio_total = float(a[0])
    else:
        audio_total = audio_s
    strong = None
    if not args.no_strong:
        try:
            strong = strong_scaling_record(args, torch, dist, runner, rank, world, local)     # every rank takes part; rank 0 gets the record
        except Exception as e:      # noqa: BLE001 -- the weak-scaling headline must still be printed
            strong = {"unavailable": f"{type(e).__name__}: {e}"}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak_tf, peak_gbs, peak_src = _peaks()
    traffic, traffic_src = None, None   # DRAM bytes per conv_umma launch from the committed ncu capture of this same command (profiles/)
    try:
        import glob
        cand = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "*_conv_umma_traffic.json")))
        if cand:
            tj = json.load(open(cand[-1]))
            traffic, traffic_src = tj["dram_bytes_per_launch"], "ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over %d launches (%s)" % (tj["launches"], os.path.basename(cand[-1]))
    except Exception:
        pass
    g = prof["conv_gemm"]
    ach = g["flops"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
    line = {
        "metric": "audio_seconds_per_second", "value": audio_total * args.steps / (dev_ms * 1e-3), "unit": "audio-s/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands, f32 accumulate / activations",
        "data": "synthetic",
        "config": {"workload": "Kokoro-82M fp16 GGUF (synthetic weights, 81.3 M params), batch 32 x 64-char (66-token) prompts per GPU, greedy durations, 24 kHz",
                   "batch_per_gpu": BATCH, "audio_s_per_step_per_gpu": audio_s, "l2": "activations (GBs per step) and weights (164 MB) exceed the 126 MB L2; no flush needed",
                   "value_timing": "CUDA events on the library stream around duration+generation passes (incl. the mid-forward host sync for durations), max over ranks; ms_per_step is this device time"},
        "e2e": {"value": audio_total * args.steps / wall, "unit": "audio-s/s", "h2d_bytes_per_step": h2d_bytes * world, "d2h_bytes_per_step": d2h_bytes * world,
                "ms_per_step": wall * 1e3 / args.steps,
                "note": "wall clock, max over ranks: host token ids in, PCM in pinned host memory out, through b2tts_kokoro_run_batch (each rank keeps its own shard's audio; the PCM leaves the device as one padded [B][S] block when padding < 25 %, d2h_bytes_per_step counts the audio itself)"},
        "gpu_launches": int(launches), "gemm_dispatch": dict(zip(("tcgen05_tma", "mma_sync_fallback"), ctx.gemm_launches())),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv_umma_kernel (persistent tcgen05 + TMA implicit-GEMM Conv1d/Linear, fp16 operands, fp32 TMEM accumulators; all conv_gemm launches incl. the few mma.sync fallbacks)", "achieved": ach, "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": g["bytes"] / max(g["launches"], 1), "algorithmic_flops_per_launch": g["flops"] / max(g["launches"], 1),
                     "peak_source": peak_src, "launches_per_step": g["launches"] / args.steps,
                     "share_of_device_time": g["ms"] / dev_ms if dev_ms else None},
        "kernel_classes_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
        "device_ms_per_step": dev_ms / args.steps,
        "pass_ms_per_step": {"duration_pass": pass_ms[0] / args.steps, "generation_pass": pass_ms[1] / args.steps},
    }
    if strong is not None:
        line["strong_scaling"] = strong
    if world == 1 and not args.no_decode_step:
        try:
            runner.close()                                     # free the Kokoro workspace (5.6 GB) before the decode model's
            line["decode_step"] = decode_step_record(ctx)
        except Exception as e:      # noqa: BLE001 -- the headline line must still be printed
            line["decode_step"] = {"unavailable": f"{type(e).__name__}: {e}"}
    if world == 1 and not args.no_cpu_baseline:
        base, why = reference_throughput(n_timed=1, n_warm=1)
        line["cpu_baseline"] = base if base else {"value": None, "unavailable": why}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
