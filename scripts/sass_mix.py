"""Summarise an `ncu --page source --csv` dump: executed warp-instructions by SASS opcode and
the hottest stall locations.  usage: sass_mix.py src.csv [topN]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); stall = []
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    n = int(r[ix["Instructions Executed"]] or 0)
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0]
    ops[op] += n; tot += n
    stall.append((int(r[ix["Warp Stall Sampling (All Samples)"]] or 0), n, src))
print("total warp-instructions", tot)
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    print("%-12s %12d  %5.1f%%" % (op, n, 100.0 * n / tot))
print("--- top stall sites")
allst = sum(s for s, _, _ in stall)
for s, n, src in sorted(stall, reverse=True)[:25]:
    print("%6d (%4.1f%%) exec=%9d  %s" % (s, 100.0 * s / allst, n, src))
