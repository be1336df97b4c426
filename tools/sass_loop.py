#!/usr/bin/env python
"""Opcode histogram of the FRAME LOOP of one sense_kernel (no GPU needed).

  python tools/sass_loop.py <obj-or-so> <mangled-name-regex> [--md]

The frame loop is found as the backward branch whose span holds the most packed-FP32 instructions; the histogram is
of the static instructions inside it - straight-line code executed once per frame and thread (both the full-frame
and the zero-padded load paths are inside, so LDG/CS2R/ISETP are counted twice).
"""
import collections
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    for f in re.split(r"\n\s+Function : ", out)[1:]:
        name = f.split("\n")[0].strip()
        ins = []
        for line in f.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        yield name, ins


def opname(text):
    toks = text.split()
    if toks[0].startswith("@"):
        toks = toks[1:]
    op = toks[0]
    base = op.split(".")[0]
    if base in ("LDG", "LDS", "STS", "STG", "LDL", "STL"):
        w = [p for p in op.split(".") if p in ("64", "128")]
        base += "." + (w[0] if w else "32")
    return base


def frame_loop(ins):
    addr = [a for a, _ in ins]
    best = None
    for i, (a, text) in enumerate(ins):
        m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", text)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= a:
            continue
        j = next(k for k, x in enumerate(addr) if x >= tgt)
        n = sum(1 for _, t in ins[j:i + 1] if re.search(r"\b(FFMA2|FADD2|FMUL2)\b", t))
        if best is None or n > best[0] or (n == best[0] and i - j < best[2] - best[1]):
            best = (n, j, i)
    return ins[best[1]:best[2] + 1]


def main():
    path, pat = sys.argv[1], re.compile(sys.argv[2])
    md = "--md" in sys.argv
    for name, ins in kernels(path):
        if not pat.search(name):
            continue
        loop = frame_loop(ins)
        ops = collections.Counter(opname(t) for _, t in loop)
        tot = sum(ops.values())
        packed = ops["FFMA2"] + ops["FADD2"] + ops["FMUL2"]
        fp32 = 2 * packed + ops["FFMA"] + ops["FADD"] + ops["FMUL"]
        lsu = sum(c for o, c in ops.items() if o.split(".")[0] in ("LDS", "STS", "LDG", "LDL", "STL"))
        print(("### `%s`\n" if md else "== %s") % name)
        print("frame loop: %d instructions, %d packed FP32 (FFMA2/FADD2/FMUL2), FP32-pipe cycles per warp %d, "
              "load/store instructions %d, local-memory (spill) %d" %
              (tot, packed, fp32, lsu, ops["LDL.32"] + ops["STL.32"] + ops["LDL.64"] + ops["STL.64"] + ops["LDL.128"] + ops["STL.128"]))
        if md:
            print("\n| opcode | count |\n|---|---|")
        for op, c in ops.most_common(24):
            print(("| %s | %d |" if md else "  %-10s %5d") % (op, c))
        print()


if __name__ == "__main__":
    main()
