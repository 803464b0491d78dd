import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CACHE = os.environ.get("B2TTS_CACHE", "/tmp/b2tts_cache")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


# GPU runs use -x: the rows that carry the headline (Kokoro, the patched ops, the codecs) are collected first, the autoregressive decode paths after them.
_LATE = ("test_orpheus_gpu", "test_parler_gpu", "test_dia_gpu", "test_sampler_gpu", "test_ar_graph_gpu", "test_ar_fullsize_gpu")


def pytest_collection_modifyitems(config, items):
    def late(item):
        mod = os.path.splitext(os.path.basename(str(item.fspath)))[0]
        return _LATE.index(mod) + 1 if mod in _LATE else 0
    items.sort(key=late)       # stable: everything else keeps its order


def synth_gguf(dtype="f16", ctx_len=128, seed=0, **kw) -> str:
    """Synthetic Kokoro GGUF, cached on disk (deterministic in its arguments)."""
    from tts_cpp_b200.synth import cached_gguf
    return cached_gguf(dtype, ctx_len, seed, cache_dir=CACHE, **kw)


@pytest.fixture(scope="session")
def gguf_path():
    return synth_gguf()


@pytest.fixture(scope="session")
def port(gguf_path):
    from oracle.kokoro_port import KokoroPort
    return KokoroPort(gguf_path)


@pytest.fixture(scope="session")
def gpu_ctx():
    from tts_cpp_b200.binding import Context
    return Context(0)


@pytest.fixture(scope="session")
def runner(gguf_path, gpu_ctx):
    from tts_cpp_b200.binding import runner_from_file
    return runner_from_file(gguf_path, ctx=gpu_ctx)


def rms(a):
    a = np.asarray(a, np.float64)
    return float(np.sqrt((a * a).mean())) if a.size else 0.0


def report(name, got, want):
    got = np.asarray(got, np.float32); want = np.asarray(want, np.float32)
    d = rms(got - want); r = rms(want)
    mx = float(np.abs(got - want).max()) if got.size else 0.0
    print(f"PARITY {name:28s} n={got.size:9d} ref_rms={r:.5g} diff_rms={d:.3g} rel={d / max(r, 1e-30):.3g} max={mx:.3g}")
    return d, r, mx


def run_snippet(code: str, args=(), env=None, timeout=300):
    """Run a test body given as source text (sys.argv[1] = repo root, then `args`); -> exit code.
    env is None: in THIS process, on the CUDA context the other tests use.  env given: in a child process -- the library reads its B2TTS_* switches once per
    process, so a variant that needs a different switch setting cannot share the parent's."""
    import subprocess
    if env is None:
        old = sys.argv
        sys.argv = ["-c", ROOT] + [str(a) for a in args]
        try:
            exec(compile(code, "<snippet>", "exec"), {"__name__": "__main__"})
            return 0
        except SystemExit as e:
            return int(e.code or 0)
        finally:
            sys.argv = old
    r = subprocess.run([sys.executable, "-c", code, ROOT] + [str(a) for a in args], capture_output=True, text=True, timeout=timeout, env=dict(os.environ, **env))
    print(r.stdout[-3000:])
    print(r.stderr[-2000:])
    return r.returncode


def tie_report(name, ref_logits, ref_tokens, got_logits, got_tokens, margin_factor=2.0):
    """Greedy-decision margins: per step the reference's smallest top-2 logit gap next to the largest logit difference of that step.  A token may differ from the
    reference only where the reference's own gap is below margin_factor x that step's max |logit difference| (a near-tie that rounding noise decides);
    -> True when every differing token is such a near-tie."""
    ref_logits = np.asarray(ref_logits, np.float32); got_logits = np.asarray(got_logits, np.float32).reshape(ref_logits.shape)
    ref_tokens = np.asarray(ref_tokens); got_tokens = np.asarray(got_tokens).reshape(ref_tokens.shape)
    ok = True
    for s in range(ref_logits.shape[0]):
        lg = ref_logits[s].reshape(ref_tokens[s].size, -1)
        top2 = np.sort(lg, axis=-1)[:, -2:]
        gap = top2[:, 1] - top2[:, 0]
        d = float(np.abs(got_logits[s] - ref_logits[s]).max())
        bad = np.nonzero(got_tokens[s].reshape(-1) != ref_tokens[s].reshape(-1))[0]
        clear = [int(h) for h in bad if gap[h] > margin_factor * d]
        print(f"TIE-MARGIN {name} step {s}: min top-2 gap {gap.min():.3e}  max |logit diff| {d:.3e}  differing heads {bad.tolist()} (gaps {[round(float(gap[h]), 4) for h in bad]})"
              + (f"  CLEAR DECISIONS DIFFER: {clear}" if clear else ""))
        ok &= not clear
    return ok
