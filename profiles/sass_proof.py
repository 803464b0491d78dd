"""profiles/sass_proof.py -- the Blackwell proof, regenerated from the library that ships (no GPU needed):
    python profiles/sass_proof.py > profiles/r2_sass_blackwell_proof.txt
`cuobjdump -sass tts_cpp_b200/libb2tts.so`, per kernel: how many tcgen05 (UTCHMMA / UTCBAR / LDTM / UTCATOMSWS = tcgen05.mma / commit / ld / alloc), TMA (UTMALDG / UTMAPF /
UBLKCP), cluster / DSMEM (STAS = st.async, UCGABAR, ACQBULK), mbarrier (SYNCS), HMMA (mma.sync) and IDP.4A (dp4a) instructions it contains, plus two verbatim SASS lines
of each tcgen05 mnemonic.  A kernel listed here with UTCHMMA > 0 issues 5th-generation tensor-core MMAs whose accumulators live in TMEM."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tts_cpp_b200", "libb2tts.so")
KEYS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "STAS", "UCGABAR", "ACQBULK", "SYNCS", "HMMA", "IMMA", "IDP", "LDGSTS")


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
    sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    print(f"# cuobjdump -sass tts_cpp_b200/libb2tts.so   (arch: {', '.join(arch)}); regenerate: python profiles/sass_proof.py")
    print("# kernel: counts of the Blackwell-specific / tensor / async SASS mnemonics (prefix match)\n")
    rows, samples = [], collections.OrderedDict()
    for fn in re.split(r"\n\s*Function : ", sass)[1:]:
        name = fn.split("\n", 1)[0].strip()
        ops = collections.Counter()
        for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)([^;]*);", fn, re.M):
            for k in KEYS:
                if m.group(1).startswith(k):
                    ops[k] += 1
                    if k in ("UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "STAS") and len(samples.setdefault(k, [])) < 2:
                        samples[k].append((m.group(1) + m.group(2)).strip())
        if ops:
            rows.append((name, ops))
    names = demangle([r[0] for r in rows])
    tot = collections.Counter()
    for n, (_, ops) in sorted(zip(names, rows)):
        print(f"{n[:110]:<112}" + "  ".join(f"{k} x{v}" for k, v in ops.items()))
        tot.update(ops)
    print("\n# totals over the library: " + ", ".join(f"{k} x{v}" for k, v in tot.items()))
    print("\n# verbatim SASS samples")
    for k, v in samples.items():
        for s in v:
            print(f"  {s}")


if __name__ == "__main__":
    main()
