#!/usr/bin/env python
"""Static opcode histogram of one kernel from `cuobjdump -sass` (no GPU needed).

  python tools/sass_hist.py cognitive-radio-network_b200/build/crn_sense_n8192.o 'HybridPlanILi8192.*Lb1ELi1ELi0ELb0ELj4294967295' [--md]

Counts are static SASS instructions of the whole function (frame loop + epilogue); the frame loop is straight-line
code executed once per frame and thread, so they are close to per-frame issue slots.  Used for profiles/*_sass_*.md.
"""
import collections
import re
import subprocess
import sys


def functions(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    name, body = None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
        elif name:
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
            if m:
                body.append(m.group(1).strip())
    if name:
        yield name, body


def hist(body):
    ops = collections.Counter()
    for ins in body:
        toks = ins.split()
        if toks[0].startswith("@"):
            toks = toks[1:]
        op = toks[0]
        base = op.split(".")[0]
        if base in ("LDG", "LDS", "STS", "STG", "LD", "ST"):
            w = [p for p in op.split(".") if p in ("64", "128", "U8", "U16", "S16")]
            base += "." + (w[0] if w else "32")
        ops[base] += 1
    return ops


def main():
    path, pat = sys.argv[1], re.compile(sys.argv[2])
    md = "--md" in sys.argv
    for name, body in functions(path):
        if not pat.search(name):
            continue
        ops = hist(body)
        tot = sum(ops.values())
        print(("### `%s`\n" if md else "== %s") % name)
        print("total %d static instructions" % tot)
        if md:
            print("\n| opcode | count | % |\n|---|---|---|")
        for op, c in ops.most_common(40 if md else 60):
            print(("| %s | %d | %.1f |" if md else "%-12s %6d %5.1f") % (op, c, 100.0 * c / tot))
        print()


if __name__ == "__main__":
    main()
