#!/usr/bin/env python
"""tools/ncu_source_mix.py SOURCE.csv -- dynamic opcode mix and stall samples of one kernel from
`ncu -i X.ncu-rep --page source --csv` (needs -lineinfo + --import-source on at capture time)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == "Address":
        hdr, start = r, i + 1
        break
ix = {n: i for i, n in enumerate(hdr)}
stallcols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
samples, insts, stalls = collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[start:]:
    if len(r) < len(hdr):
        continue
    toks = r[ix["Source"]].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    samples[op] += int(r[ix["# Samples"]] or 0)
    insts[op] += int(r[ix["Instructions Executed"]] or 0)
    for c in stallcols:
        stalls[c] += int(r[ix[c]] or 0)
ts, ti = sum(samples.values()), sum(insts.values())
print(f"warp instructions executed {ti}, stall samples {ts}")
for op, n in insts.most_common(16):
    print(f"{op:10s} inst {n:11d} {100 * n / ti:5.1f}%   samples {samples[op]:7d} {100 * samples[op] / max(ts, 1):5.1f}%")
print({k: v for k, v in stalls.most_common(8)})
