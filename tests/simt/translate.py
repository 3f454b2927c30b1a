#!/usr/bin/env python3
"""simt-check: rewrites the two CUDA-only constructs of a .cu/.cuh source so that g++ can
compile it against tests/simt/simt.h (test infrastructure only):

  kernel<<<grid, block, smem, stream>>>(args);   ->  simt::launch(grid, block, smem, [=]() { kernel(args); });
  extern __shared__ ... name[];                   ->  unsigned char *name = simt::dyn_smem();

usage: translate.py in.cu out.cpp
"""
import re
import sys


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def translate(src):
    out = []
    pos = 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            out.append(src[pos:])
            break
        # kernel name: identifier (with optional template arguments) right before <<<
        j = i
        if src[j - 1] == ">":
            depth = 0
            while True:
                j -= 1
                if src[j] == ">":
                    depth += 1
                elif src[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
        while j > 0 and (src[j - 1].isalnum() or src[j - 1] == "_"):
            j -= 1
        name = src[j:i]
        k = src.index(">>>", i)
        cfg = split_top(src[i + 3:k])
        while len(cfg) < 3:
            cfg.append("0")
        # argument list
        a = src.index("(", k)
        depth, e = 0, a
        while True:
            if src[e] == "(":
                depth += 1
            elif src[e] == ")":
                depth -= 1
                if depth == 0:
                    break
            e += 1
        args = src[a + 1:e]
        out.append(src[pos:j])
        out.append("simt::launch(%s, %s, %s, [=]() { %s(%s); })" % (cfg[0], cfg[1], cfg[2], name, args))
        pos = e + 1
    text = "".join(out)
    text = re.sub(r"extern\s+__shared__[^;]*?(\w+)\s*\[\s*\]\s*;", r"unsigned char *\1 = simt::dyn_smem();", text)
    return text


if __name__ == "__main__":
    with open(sys.argv[1]) as f:
        src = f.read()
    with open(sys.argv[2], "w") as f:
        f.write('#line 1 "%s"\n' % sys.argv[1])
        f.write(translate(src))
