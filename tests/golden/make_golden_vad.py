#!/usr/bin/env python3
"""tests/golden/make_golden_vad.py -- runs the compiled UNMODIFIED reference (oracle/_ref/vad_ref = examples/cli/vad.cpp + oracle/ref_vad_driver.cpp) on the cases of
vad_cases.py and stores what it returned: tests/golden/vad_vectors.npz (trimmed lengths, frame energies).  Only runs where oracle/_ref exists."""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
import vad_cases  # noqa: E402


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "vad_ref")
    out = {}
    for name, kw, utts in vad_cases.cases():
        with tempfile.TemporaryDirectory() as tmp:
            fi, fo = os.path.join(tmp, "i.bin"), os.path.join(tmp, "o.bin")
            open(fi, "wb").write(vad_cases.pack_input(kw, utts))
            subprocess.run([exe, fi, fo], check=True)
            n_out, en = vad_cases.unpack_output(open(fo, "rb").read(), kw, utts)
        out[name + ".n_in"] = np.asarray([u.size for u in utts], np.int64)
        out[name + ".n_out"] = n_out
        for b, e in enumerate(en):
            out[f"{name}.energies.{b}"] = e
        print(name, [u.size for u in utts], "->", n_out.tolist())
    np.savez_compressed(os.path.join(HERE, "vad_vectors.npz"), **out)


if __name__ == "__main__":
    main()
