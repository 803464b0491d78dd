"""GPU: the on-device sampler (tts_cpp_b200/csrc/sampler.cu) through the C-ABI (b2tts_op_sample) against oracle/sampler_port.py -- itself pinned to the
reference sampler stage by stage and by a histogram of its draws (tests/test_oracle_port.py).  The port is fed the uniforms the kernel derives from
(seed, row, step): tokens and repetition state must be identical over consecutive steps.

Passed on a B200 (GPUTEST_r01, gpurun_out/r2a), including the 156 940-wide rows: a plain in-process test."""
import pytest

from conftest import run_snippet

pytestmark = pytest.mark.gpu

CHILD = r'''
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from tts_cpp_b200.binding import Context, lib, _chk, sample_uniform
from oracle.sampler_port import SamplerPort, uniform_from_counter
cfgs = {"greedy": (0, 1.0, 0, 1.0, 1.0), "default_top50": (1, 1.0, 50, 1.0, 1.0), "temp_rep": (1, 0.7, 20, 1.0, 1.3), "topk_topp": (1, 1.3, 40, 0.9, 1.0),
        "topp_only": (1, 0.9, 0, 0.8, 1.1), "full_vocab": (1, 1.1, 0, 1.0, 1.2)}
ctx = Context(0)
rng = np.random.default_rng(33)
seed = 0x1234ABCD5678
if sys.argv[2] == "wide":       # Orpheus' vocabulary on a coarse grid of values: the nucleus boundary falls inside a run of equal logits (the radix select's ordered tie path)
    rows, V, steps = 6, 156940, 3
    logits = (np.round(rng.standard_normal((steps, rows, V)) * 4.0) / 2.0).astype(np.float32)
    logits[:, :, 11] += 3.0
    cfgs = {k: cfgs[k] for k in ("default_top50", "temp_rep", "topk_topp", "topp_only")}      # topp_only: a nucleus of tens of thousands of entries, produced in chunks of 1 024
    cfgs["top1000_flat"] = (1, 4.0, 1000, 1.0, 1.0)
else:
    rows, V, steps = 18, 1088, 6
    logits = (rng.standard_normal((steps, rows, V)) * 2.5).astype(np.float32)
    logits[:, :, 7] += 6.0
    logits[2:, 1, 40] = logits[2:, 1, 41]
ok = True
for name, (do_sample, temp, top_k, top_p, rp) in cfgs.items():
    port = SamplerPort(rows, V, temp, top_k, top_p, rp)
    last = np.full(rows, -1, np.int32); counts = np.zeros(rows, np.int32)
    for s in range(steps):
        toks = np.empty(rows, np.int32)
        _chk(lib().b2tts_op_sample(ctx.h, logits[s].ctypes.data_as(C.POINTER(C.c_float)), rows, V, do_sample, top_k, C.c_float(top_p), C.c_float(temp), C.c_float(rp),
                                   last.ctypes.data_as(C.POINTER(C.c_int32)), counts.ctypes.data_as(C.POINTER(C.c_int32)), C.c_uint64(seed), s,
                                   toks.ctypes.data_as(C.POINTER(C.c_int32))))
        us = np.array([uniform_from_counter(seed, r, s) for r in range(rows)], np.float32)
        assert all(sample_uniform(seed, r, s) == us[r] for r in range(rows))
        want = port.draw(logits[s], us) if do_sample else np.array([int(np.argmax(port._eff(logits[s][i], i))) for i in range(rows)])
        good = bool(np.array_equal(toks, want))
        ok &= good
        if not good: print(f"MISMATCH {name} step {s}: {toks.tolist()} vs {want.tolist()}")
    if do_sample and rp != 1.0:
        ok &= bool(np.array_equal(last, port.last)) and bool(np.array_equal(counts, port.counts))
    print(f"PARITY sampler {name}: {'ok' if ok else 'FAILED'}")
sys.exit(0 if ok else 1)
'''


@pytest.mark.parametrize("shape", ["small", "wide"])
def test_sampler_matches_port_over_steps(shape):
    assert run_snippet(CHILD, [shape]) == 0
