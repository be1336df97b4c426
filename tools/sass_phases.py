#!/usr/bin/env python
"""Group an `ncu --page source --csv --print-source sass` dump by opcode class and list the hottest stall sites.
   ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv;  python tools/sass_phases.py src.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, ismp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ops = collections.Counter()
smp = collections.Counter()
tot_ex = tot_smp = 0
body = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[isrc].strip()
    toks = src.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    base = op.split(".")[0]
    ex, s = int(r[iex]), int(r[ismp])
    ops[base] += ex
    smp[base] += s
    tot_ex += ex
    tot_smp += s
    body.append((s, ex, src, {hdr[i]: int(r[i]) for i in stall_cols if int(r[i])}))
print("total warp instructions %d, samples %d" % (tot_ex, tot_smp))
print("%-12s %14s %6s %8s %6s" % ("opcode", "warp instr", "%", "samples", "%"))
for op, ex in ops.most_common(30):
    print("%-12s %14d %6.1f %8d %6.1f" % (op, ex, 100.0 * ex / tot_ex, smp[op], 100.0 * smp[op] / max(tot_smp, 1)))
print("\nhottest sampling sites:")
for s, ex, src, st in sorted(body, key=lambda b: -b[0])[:25]:
    print("%6d %10d  %-60s %s" % (s, ex, src[:60], dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])))
