"""profiles/static_ar_summary.py -- regenerate the STATIC evidence for the autoregressive decode kernels (no GPU needed):
    python profiles/static_ar_summary.py > profiles/r1h_ar_kernels_ptxas.txt
ptxas -v (registers / stack / spills) for every kernel of parler.cu's copy of ar_kernels.cuh and of sampler.cu, and SASS mnemonic counts (cuobjdump -sass) of the
kernels the decode roofline is about.  Static only: none of these kernels has run on a B200 yet."""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "tts_cpp_b200", "csrc")
NVCC = "/usr/local/cuda/bin/nvcc"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--fmad=false", "-Xptxas", "-v"]


def demangle(names):
    r = subprocess.run(["/usr/local/cuda/bin/cu++filt"] + names, capture_output=True, text=True)
    out = []
    for line in r.stdout.splitlines():
        line = re.sub(r"b2::(\(anonymous namespace\)|<unnamed>)::", "", line)
        line = re.sub(r"^void ", "", line)
        i = line.rfind(">(")
        out.append(line[:i + 1] if i >= 0 else re.sub(r"\(.*$", "", line))
    return out


def main():
    tmp = tempfile.mkdtemp(prefix="b2static_")
    print("# ptxas -v for the autoregressive decode kernels as compiled into libb2tts.so (parler.cu's copy of ar_kernels.cuh, sampler.cu); sm_100a, -O3, --fmad=false")
    print("# (static evidence only: these kernels have not run on a GPU yet -- DESIGN.md 7.1).  Regenerate: python profiles/static_ar_summary.py\n")
    objs = {}
    for src in ("parler.cu", "sampler.cu"):
        obj = os.path.join(tmp, src + ".o")
        r = subprocess.run([NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        if r.returncode:
            sys.exit(r.stderr)
        objs[src] = obj
        ents = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used \d+ barriers)?(?:, \d+ bytes cumulative stack size)?(?:, (\d+) bytes smem)?", r.stderr)
        names = demangle([e[0] for e in ents])
        for n, e in sorted(zip(names, ents)):
            print(f"{n:<52} {e[4]:>3} regs   stack {e[1]} B, spills {e[2]}/{e[3]} B" + (f", static smem {e[5]} B" if e[5] else ""))
    print()
    sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", objs["parler.cu"]], capture_output=True, text=True).stdout
    sass += subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", objs["sampler.cu"]], capture_output=True, text=True).stdout
    want = ["gemv_mma_group_kernelILb0ELi1", "gemv_mma_group_kernelILb1ELi4", "gemv_rows_q_group_kernel", "gemv_rows_group_kernelI6__halfLb1ELi2", "attention_gqa_kernel", "sample_rows_kernel"]
    keep = re.compile(r"^(HMMA|IDP|LDG|LDS|STS|ATOMS|ATOM|RED|MUFU|FFMA|DADD|DFMA|BAR|SHFL|POPC)")
    for fn in re.split(r"\n\s*Function : ", sass)[1:]:
        name = fn.split("\n", 1)[0].strip()
        if not any(w in name for w in want):
            continue
        ops = collections.Counter()
        for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", fn, re.M):
            if keep.match(m.group(1)):
                ops[m.group(1)] += 1
        print(f"{demangle([name])[0]}: SASS mnemonic counts: " + ", ".join(f"{k} x{v}" for k, v in sorted(ops.items())))


if __name__ == "__main__":
    main()
