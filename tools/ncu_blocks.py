#!/usr/bin/env python
"""tools/ncu_blocks.py SOURCE.csv -- group the SASS rows of `ncu --page source --csv` into runs of equal execution
count (~ basic blocks / loop levels) and print, per run, its share of the executed warp instructions, of the stall
samples, and its opcode histogram.  Shows at a glance where the non-MAC instructions of a kernel live."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == "Address":
        hdr, start = r, i + 1
        break
ix = {n: i for i, n in enumerate(hdr)}
blocks, cur = [], None
for k, r in enumerate(rows[start:]):
    if len(r) < len(hdr):
        continue
    toks = r[ix["Source"]].split()
    if not toks:
        continue
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    n, s = int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0)
    if cur is None or cur["n"] != n:
        cur = {"n": n, "first": k, "ops": collections.Counter(), "samples": 0, "len": 0}
        blocks.append(cur)
    cur["ops"][op] += 1
    cur["samples"] += s
    cur["len"] += 1
ti = sum(b["n"] * b["len"] for b in blocks)
ts = sum(b["samples"] for b in blocks)
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
print(f"warp instructions {ti}, samples {ts}")
for b in blocks:
    share = 100.0 * b["n"] * b["len"] / ti
    if share < thr and 100.0 * b["samples"] / max(ts, 1) < thr:
        continue
    ops = " ".join(f"{o}:{c}" for o, c in b["ops"].most_common(9))
    print(f"row {b['first']:5d} len {b['len']:4d} x {b['n']:9d}  inst {share:5.1f}%  samples {100.0 * b['samples'] / max(ts, 1):5.1f}%  {ops}")
