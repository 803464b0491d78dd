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


# GPU runs use -x: the rows that carry the headline (Kokoro, the patched ops, the codecs) are collected first, the autoregressive decode paths (most of whose
# variants have not run on hardware yet) after them, so that a failure there cannot keep the headline's parity tests from running.
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
