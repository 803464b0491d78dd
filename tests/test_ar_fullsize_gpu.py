"""GPU: the Parler decode loop at BASELINE config 3's FULL model size (Parler-TTS-Mini-shaped F16 decoder: 24 layers x 1024, 16 heads, ffn 4096, nine 1088-wide heads;
synthetic weights).  A golden from the compiled reference at this size pins the first steps (tests/golden/parler_mini_vectors.npz, made by tests/golden/make_golden.py:
two prompts, 24 teacher-forceable frames with logits); size-independent properties cover the rest:

  * the persistent kernel's stream is deterministic: a sequence's tokens do not depend on what else is in the batch (ragged prompts), a shorter run is a prefix of a
    longer one (launch boundaries every 32 steps included), and the fp32-page variant (B2TTS_KV=f32) agrees wherever decisions are clear;
  * against the launch-per-op path (B2TTS_AR_PDK=0: tensor-core GEMV + CUDA-graph replay, contiguous fp32 cache), teacher-forced on the persistent kernel's own tokens:
    logits within the F16 floor and the same token wherever the top-2 gap exceeds twice the step's largest logit difference;
  * within the launch-per-op family: fused = unfused (B2TTS_AR_FUSE=0) and graph replay = direct launches, token for token (same arithmetic per output)."""
import os

import numpy as np
import pytest

from conftest import run_snippet

pytestmark = pytest.mark.gpu

BODY = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from tts_cpp_b200.binding import parler_runner_from_file
from tts_cpp_b200.synth import PARLER_MINI_SHAPE, cached_parler_gguf
out, steps, which = sys.argv[2], int(sys.argv[3]), sys.argv[4]
par = parler_runner_from_file(cached_parler_gguf(seed=0, f16=True, **PARLER_MINI_SHAPE))
rng = np.random.default_rng(11)
prompts = [rng.integers(1, 500, size=n).astype(np.uint32) for n in (24, 9, 17, 31)]
if which == "teacher":
    teacher = np.load(sys.argv[5])
    toks, logits = par.generate_teacher_forced(prompts, teacher)
    np.savez(out, toks=toks, logits=logits)
else:
    toks = par.generate_greedy(prompts if which == "batch" else [prompts[2]], steps)
    np.savez(out, toks=np.stack([np.asarray(t) for t in toks]))
print("decode ms", par.last_ms(), "persistent-kernel launches / steps", par.pdk_stats())
par.close()
'''


def _run(tmp_path, tag, steps, which="batch", env=None, teacher=None):
    out = str(tmp_path / f"{tag}.npz")
    assert run_snippet(BODY, [out, steps, which] + ([teacher] if teacher else []), env=env, timeout=600) == 0
    return np.load(out)


def test_parler_mini_size_persistent_kernel_properties(tmp_path):
    steps = 40
    base = _run(tmp_path, "pdk", steps)["toks"]                            # [4][steps][9], two launches of the persistent kernel (32 + 8 steps)
    assert base.shape == (4, steps, 9) and base.min() >= 0 and base.max() < 1088
    assert len(np.unique(base)) > 50                                       # not a degenerate constant stream
    alone = _run(tmp_path, "alone", steps, which="single")["toks"]
    assert np.array_equal(alone[0], base[2])                               # batching does not change a sequence
    short = _run(tmp_path, "short", 12)["toks"]
    assert np.array_equal(short, base[:, :12])                             # a shorter run is a prefix of a longer one
    # teacher-forced on the persistent kernel's tokens: persistent (fp16 pages), persistent (fp32 pages), launch-per-op
    tf = str(tmp_path / "teacher.npy")
    np.save(tf, base)
    a = _run(tmp_path, "tf_pdk", steps, which="teacher", teacher=tf)
    assert np.array_equal(a["toks"], base)                                 # feeding back its own tokens reproduces them
    for tag, env in (("tf_pdk_kv_f32", {"B2TTS_KV": "f32"}), ("tf_ops", {"B2TTS_AR_PDK": "0"})):
        b = _run(tmp_path, tag, steps, which="teacher", env=env, teacher=tf)
        d = np.abs(b["logits"] - a["logits"])
        rms = float(np.sqrt((d.astype(np.float64) ** 2).mean()))
        top2 = np.sort(a["logits"], axis=-1)[..., -2:]
        gap = top2[..., 1] - top2[..., 0]                                  # [4][steps][9]
        dmax = d.max(axis=(2, 3))[..., None]                               # per (sequence, step)
        clear = gap > 2.0 * dmax
        agree = float((b["toks"] == base).mean())
        print(f"PARITY parler-mini {tag} vs persistent kernel (teacher-forced): logit diff rms {rms:.3e} max {float(d.max()):.3e}; tokens equal {agree:.4f}; clear decisions {float(clear.mean()):.3f}")
        assert rms < 2e-2 and float(d.max()) < 0.3
        assert np.array_equal(b["toks"][clear], base[clear])


def test_parler_mini_size_launch_per_op_family_is_self_consistent(tmp_path):
    steps = 24
    ops = _run(tmp_path, "ops", steps, env={"B2TTS_AR_PDK": "0"})["toks"]
    unfused = _run(tmp_path, "unfused", steps, env={"B2TTS_AR_PDK": "0", "B2TTS_AR_FUSE": "0"})["toks"]
    assert np.array_equal(ops, unfused)
    direct = _run(tmp_path, "direct", steps, env={"B2TTS_AR_PDK": "0", "B2TTS_AR_GRAPH": "0"})["toks"]
    assert np.array_equal(ops, direct)


GOLD_BODY = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from conftest import tie_report
from tts_cpp_b200.binding import parler_runner_from_file
from tts_cpp_b200.synth import PARLER_MINI_SHAPE, cached_parler_gguf
g = np.load(os.path.join(sys.argv[1], "tests", "golden", "parler_mini_vectors.npz"))
par = parler_runner_from_file(cached_parler_gguf(seed=0, f16=True, **PARLER_MINI_SHAPE))
prompts = [g["prompt0"], g["prompt1"]]
steps = g["tokens0"].shape[0]
toks, logits = par.generate_greedy(prompts, steps, want_logits=True)
tf_t, tf_l = par.generate_teacher_forced(prompts, np.stack([g["tokens0"], g["tokens1"]]))
ok = True
for u in range(2):
    rl, rt = g[f"logits{u}"], g[f"tokens{u}"]
    rms = np.sqrt(((tf_l[u] - rl).astype(np.float64) ** 2).mean(axis=(1, 2)))
    mx = np.abs(tf_l[u] - rl).max(axis=(1, 2))
    print(f"PARITY parler-mini f16 prompt {u} teacher-forced vs reference: logit rms max {rms.max():.3e}  max |d| {mx.max():.3e} (logit std {rl.std():.2f}); tokens equal {int((tf_t[u] == rt).sum())}/{rt.size}")
    ok &= float(rms.max()) < 1e-2 and float(mx.max()) < 0.1
    ok &= tie_report(f"parler-mini prompt {u} (teacher-forced)", rl, rt, tf_l[u], tf_t[u])
    top2 = np.sort(rl, axis=-1)[..., -2:]
    gap = (top2[..., 1] - top2[..., 0]).min(axis=1)
    first_tie = next((s for s in range(steps) if gap[s] <= 2.0 * mx[s]), steps)
    same = bool(np.array_equal(toks[u][:first_tie], rt[:first_tie]))
    print(f"PARITY parler-mini f16 prompt {u} free-running: identical through step {first_tie - 1}: {same}; all {steps} steps equal: {bool(np.array_equal(toks[u], rt))}")
    ok &= same
print("persistent-kernel launches / steps", par.pdk_stats())
ok &= (par.pdk_stats()[1] == 2 * steps) == (os.environ.get("B2TTS_AR_PDK") != "0")
par.close()
sys.exit(0 if ok else 1)
'''


@pytest.mark.parametrize("path", ["persistent_kernel", "launch_per_op"])
def test_parler_mini_size_matches_reference(path):
    """BASELINE config 3's model size against the compiled unmodified reference (tests/golden/parler_mini_vectors.npz: 24 frames, two prompts): teacher-forced logits
    within 1e-2 RMS / 0.1 max of the reference's, the same token wherever the reference's decision is clear, and free-running identical up to the first near-tie."""
    assert run_snippet(GOLD_BODY, [], env=None if path == "persistent_kernel" else {"B2TTS_AR_PDK": "0"}, timeout=600) == 0
