#!/usr/bin/env python
"""Joins an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) with nvdisasm
--print-line-info output of the same cubin, to get instructions executed / stall samples
per CUDA source line.  usage: sass_by_line.py src.csv disasm.txt <mangled kernel name> [topN]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, disasm, kname = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 60
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ic, ismp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
ith = hdr.index("Thread Instructions Executed")
sass = [(r[isrc].strip(), int(r[ic]), int(r[ismp]), int(r[ith])) for r in rows[2:] if len(r) > ic and r[ic].isdigit()]

lines = open(disasm).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kname + ":"))
cur = None
dis = []
stack = []
for l in lines[start + 1:]:
    if l.startswith("//-----") or l.startswith("\t.section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), m.group(3))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        dis.append((m.group(2).strip(), cur))
print("sass rows", len(sass), "disasm instrs", len(dis))
per = defaultdict(lambda: [0, 0, 0])
tot = sum(s[1] for s in sass)
n = min(len(sass), len(dis))
mism = 0
for i in range(n):
    op_a = sass[i][0].split()[0].lstrip("@!P0123456789UT ") if sass[i][0] else ""
    if sass[i][0].split()[-1][:3] != dis[i][0].split()[-1][:3]:
        mism += 1
    key = dis[i][1][:2] if dis[i][1] else ("?", 0)
    # inlined-at chains: attribute to the innermost line (what nvdisasm prints last)
    per[key][0] += sass[i][1]; per[key][1] += sass[i][2]; per[key][2] += sass[i][3]
print("opcode mismatches", mism, "total inst", tot)
items = sorted(per.items(), key=lambda kv: -kv[1][0])
srcs = {}
for (f, ln), (n_i, smp, th) in items[:topn]:
    if f not in srcs:
        try:
            srcs[f] = open("/root/repo/wfa_b200/csrc/" + f).read().split("\n")
        except Exception:
            srcs[f] = []
    text = srcs[f][ln - 1].strip()[:110] if 0 < ln <= len(srcs[f]) else ""
    print("%6.2f%% inst  smp %6d  thr/inst %4.1f  %s:%d  %s" % (100.0 * n_i / tot, smp, th / max(n_i, 1), f, ln, text))
