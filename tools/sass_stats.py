#!/usr/bin/env python3
"""Per-kernel SASS instruction mix of a cubin/.so (uses cuobjdump).  Usage: sass_stats.py lib.so [filter]"""
import subprocess, sys, re, collections
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
flt = sys.argv[2] if len(sys.argv) > 2 else ""
cur = None; stats = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: cur = m.group(1); stats[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+(\.[A-Z0-9_.]+)?)", line)
    if m and cur:
        op = m.group(2)
        base = op.split(".")[0]
        if base == "IMAD" and ".WIDE" in op: base = "IMAD.WIDE"
        stats[cur][base] += 1
for k, c in stats.items():
    if flt not in k: continue
    tot = sum(c.values())
    print(f"{k}: {tot} instrs; " + ", ".join(f"{o}={n}" for o, n in c.most_common(14)))
