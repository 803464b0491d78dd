"""tests/emu/emu_build.py -- TEST INFRASTRUCTURE ONLY: compile unmodified .cu files of the library against the CPU emulation (tests/emu/include)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "tts_cpp_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
sys.path.insert(0, HERE)
import prep  # noqa: E402


def build(name, cu_sources, cpp_sources, defines=(), asan=False):
    """-> path of the executable; rebuilt when any input is newer"""
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, name)
    srcs = []
    for h in os.listdir(CSRC):                      # headers with device code are converted too; BUILD is searched first (same directory as the converted sources)
        if h.endswith(".cuh"):
            text = open(os.path.join(CSRC, h)).read()
            if "<<<" in text or "extern __shared__" in text:
                open(os.path.join(BUILD, h), "w").write('#line 1 "%s"\n' % os.path.join(CSRC, h) + prep.convert(text))
    for cu in cu_sources:
        out = os.path.join(BUILD, os.path.basename(cu)[:-3] + ".cpp")
        open(out, "w").write('#line 1 "%s"\n' % os.path.join(CSRC, cu) + prep.convert(open(os.path.join(CSRC, cu)).read()))
        srcs.append(out)
    srcs += [os.path.join(HERE, "emu.cpp")] + [os.path.join(HERE, c) if not os.path.isabs(c) else c for c in cpp_sources]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "include", f) for f in os.listdir(os.path.join(HERE, "include"))]
    if os.path.exists(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(d) for d in deps if not d.startswith(BUILD)):
        return exe
    opt = ["-O1", "-g", "-fsanitize=address,alignment", "-fno-sanitize-recover=alignment", "-fno-omit-frame-pointer"] if asan else ["-O3", "-march=native", "-fno-plt"]
    cmd = ["g++"] + opt + ["-std=c++17", "-DB2EMU", "-I" + os.path.join(HERE, "include"), "-I" + BUILD, "-I" + CSRC] + ["-D" + d for d in defines] + ["-o", exe] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("emulation build failed:\n" + r.stderr[-4000:])
    return exe
