"""tests/emu/prep.py -- TEST INFRASTRUCTURE ONLY: rewrite the two pieces of CUDA syntax a host compiler cannot parse, so that an unmodified
.cu file of the library compiles against tests/emu/include:

    kernel<<<grid, block, smem, stream>>>(args);   ->  b2emu::launch(grid, block, smem, [=]() { kernel(args); });     (by value, like kernel parameters:
                                                    a launch recorded during stream capture can be replayed after the caller's locals are gone)
    extern __shared__ T name[];                    ->  T * name = (T *) b2emu::dyn_smem();
"""
import re
import sys


def _match(src, i, open_c, close_c):
    depth = 0
    while i < len(src):
        c = src[i]
        if c == open_c:
            depth += 1
        elif c == close_c:
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced")


def _split_top(s):
    out, depth, cur = [], 0, ""
    for c in s:
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += c
    out.append(cur.strip())
    return out


def convert(src):
    src = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];", r"\1 * \2 = (\1 *) b2emu::dyn_smem();", src)
    out, pos = "", 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            break
        m = re.search(r"([A-Za-z_][\w:]*(?:<[^<>;(){}]*>)?)\s*$", src[pos:i])
        if not m:
            raise ValueError("no kernel name before <<< at %d" % i)
        kstart = pos + m.start(1)
        j = src.index(">>>", i)
        cfg = _split_top(src[i + 3:j])
        k = j + 3
        while src[k].isspace():
            k += 1
        assert src[k] == "(", src[k:k + 40]
        e = _match(src, k, "(", ")")
        grid, block = cfg[0], cfg[1]
        smem = cfg[2] if len(cfg) > 2 else "0"
        out += src[pos:kstart] + "b2emu::launch(%s, %s, %s, [=]() { %s%s; })" % (grid, block, smem, m.group(1), src[k:e + 1])
        pos = e + 1
    return out + src[pos:]


if __name__ == "__main__":
    open(sys.argv[2], "w").write(convert(open(sys.argv[1]).read()))
