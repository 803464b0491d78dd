"""Build libb2tts.so (the sm_100a CUDA kernels + C-ABI) in-tree with nvcc.

    python -m tts_cpp_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU.  The .so lands next to this file
(git-ignored, but shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libb2tts.so")
OBJDIR = os.path.join(HERE, "build")
SOURCES = ["capi.cu", "kokoro.cu", "dac.cu", "orpheus.cu", "parler.cu", "dia.cu", "t5.cu", "pdk.cu", "gemm_conv.cu", "gemm_umma.cu", "lstm.cu", "elementwise.cu", "source.cu", "sampler.cu", "vad.cu", "gguf_reader.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3", "--fmad=false",   # parity: no implicit a*b+c contraction; hot loops call fmaf explicitly
    "-Xptxas", "-v" if os.environ.get("B2TTS_PTXAS_V") else "-O3",
]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "b2tts.h"))
    objs, jobs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJDIR, s + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [NVCC] + FLAGS + (["-x", "cu"] if s.endswith(".cpp") else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose and os.environ.get("B2TTS_PTXAS_V"):
            sys.stderr.write(r.stderr)
        return 0

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if force or jobs or _stale(OUT, objs):
        run([NVCC, "-shared", "-o", OUT] + objs + ["-lcudart"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
