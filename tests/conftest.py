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
_LATE = ("test_orpheus_gpu", "test_parler_gpu", "test_dia_gpu", "test_sampler_gpu", "test_ar_graph_gpu", "test_ar_fullsize_gpu", "test_vad_gpu", "test_server_gpu", "test_t5_gpu")


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


class patched_server:
    """integration/_build/tts-server-b200 (the reference's server.cpp with the batch-draining worker patched in at build time) on 127.0.0.1.
    `with patched_server(model_path) as url:`; the process is ended by PID.  `.log()` = its stderr (B2TTS_WORKER_LOG=1: one line per batched forward)."""

    def __init__(self, model_path, max_batch=32, extra=()):
        self.exe = os.path.join(ROOT, "integration", "_build", "tts-server-b200")
        self.model_path, self.max_batch, self.extra = model_path, max_batch, list(extra)

    def __enter__(self):
        import socket
        import subprocess
        import tempfile
        import time
        import requests
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        self.errf = tempfile.NamedTemporaryFile(prefix="b2srv_", suffix=".log", delete=False)
        env = dict(os.environ, B2TTS_WORKER_LOG="1", B2TTS_SERVER_MAX_BATCH=str(self.max_batch))
        self.p = subprocess.Popen([self.exe, "--model-path", self.model_path, "--port", str(port), "--n-threads", "4"] + self.extra, stdout=subprocess.DEVNULL,
                                  stderr=self.errf, env=env, cwd=tempfile.gettempdir())
        self.url = f"http://127.0.0.1:{port}"
        for _ in range(600):                                   # model load + CUDA context on a fresh box can take a while
            if self.p.poll() is not None:
                raise RuntimeError("server exited: " + self.log()[-2000:])
            try:
                if requests.get(self.url + "/health", timeout=1).status_code == 200:
                    return self
            except requests.RequestException:
                pass
            time.sleep(0.25)
        self.__exit__(None, None, None)
        raise RuntimeError("server did not come up: " + self.log()[-2000:])

    def speech(self, prompts, threads=8, **fields):
        """POST /v1/audio/speech for every prompt, `threads` at a time -> list of (status, int16 samples, frame rate)."""
        import concurrent.futures as cf
        import io
        import wave
        import requests

        def one(p):
            r = requests.post(self.url + "/v1/audio/speech", json=dict(input=p, **fields), timeout=300)
            if r.status_code != 200:
                return r.status_code, np.zeros(0, np.int16), 0
            with wave.open(io.BytesIO(r.content), "rb") as w:
                return 200, np.frombuffer(w.readframes(w.getnframes()), np.int16), w.getframerate()
        with cf.ThreadPoolExecutor(threads) as ex:
            return list(ex.map(one, prompts))

    def log(self):
        self.errf.flush()
        return open(self.errf.name, errors="replace").read()

    def forwards(self):
        import re
        return [int(m) for m in re.findall(r"b200 worker: forward of (\d+) task", self.log())]

    def __exit__(self, *a):
        import subprocess
        if self.p.poll() is None:
            self.p.terminate()
            try:
                self.p.wait(timeout=10)
            except subprocess.TimeoutExpired:
                self.p.kill()
                self.p.wait()
