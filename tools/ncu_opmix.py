#!/usr/bin/env python3
"""Dynamic (executed) SASS opcode mix from an `ncu --page source --csv` export.
Usage: ncu -i rep.ncu-rep --page source --csv > src.csv ; ncu_opmix.py src.csv"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
mix = collections.Counter(); samp = collections.Counter(); tot = 0
for r in rows[2:]:
    if len(r) <= iex: continue
    src = r[isrc].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    if not m: continue
    op = m.group(2)
    base = op.split(".")[0]
    if base == "IMAD":
        base = "IMAD.WIDE" if ".WIDE" in op else ("IMAD.MOV" if ".MOV" in op else ("IMAD.X" if ".X" in op else ("IMAD.IADD" if ".IADD" in op else ("IMAD.HI" if ".HI" in op else "IMAD"))))
    if base == "IADD3" and ".X" in op: base = "IADD3.X"
    n = int(r[iex] or 0); mix[base] += n; tot += n; samp[base] += int(r[isamp] or 0)
print("total warp-instructions executed:", tot)
for k, v in mix.most_common(24):
    print(f"{k:12s} {v:12d} {100*v/tot:6.2f}%   samples {samp[k]}")
