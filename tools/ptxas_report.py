#!/usr/bin/env python
"""Registers / spills per sense_kernel instantiation: python tools/ptxas_report.py [size ...] [-D...]
Compiles csrc/crn_sense_n<size>.cu with -Xptxas -v (no GPU needed) and prints one line per kernel."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cognitive-radio-network_b200")]
import build as B  # noqa: E402


def report(size, defs):
    src = os.path.join(B.CSRC, "crn_sense_n%d.cu" % size)
    cmd = [B._nvcc()] + B.ARCH + B.NVCC_FLAGS + defs + ["-Xptxas", "-v", "-x", "cu", "-c", src, "-o", "/tmp/ptxas_report_%d.o" % size]
    err = subprocess.run(cmd, capture_output=True, text=True).stderr
    name = None
    rows = []
    for line in err.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = m.group(1)
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            spill = (int(m.group(2)), int(m.group(3)))
        m = re.search(r"Used (\d+) registers", line)
        if m and name:
            t = re.search(r"EEELb(\d)ELi(\d)ELi(\d)ELb(\d)ELj(\d+)E", name)
            tag = "win=%s det=%s epi=%s sc16=%s %s" % (t.group(1), "magsq" if t.group(2) == "1" else "mag",
                                                       "cta" if t.group(3) == "0" else "unit", t.group(4),
                                                       "all" if t.group(5) == "4294967295" or t.group(5) == "65535" else "ref")
            rows.append((tag, int(m.group(1)), spill))
            name = None
    for tag, regs, spill in sorted(rows):
        print("N=%-5d %-42s regs %3d  spill st/ld %3d/%3d" % (size, tag, regs, spill[0], spill[1]))


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [256, 512, 1024, 2048, 4096, 8192]
    defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    for n in sizes:
        report(n, defs)
