"""Static size of the main loop of a kernel in a .so: finds the last backward branch and reports
instruction counts by opcode between its target and itself.  usage: loop_size.py lib.so <mangled substring>"""
import subprocess, sys, re, collections
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
cur = None; funcs = {}
for l in out.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m: cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and cur: funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if sys.argv[2] not in name: continue
    best = None
    for a, t in ins:
        m = re.search(r"BRA\S*\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
        if m and m.group(1):
            tgt = int(m.group(1), 16)
            if tgt < a and (best is None or a - tgt > best[1] - best[0]): best = (tgt, a)
    print(name[:80], "total", len(ins))
    if best:
        body = [t for a, t in ins if best[0] <= a <= best[1]]
        c = collections.Counter()
        for t in body:
            tk = t.split(); o = tk[1] if tk[0].startswith("@") else tk[0]
            c[o.split(".")[0]] += 1
        print(" loop", hex(best[0]), hex(best[1]), "instructions", len(body))
        print(" ", ", ".join("%s %d" % kv for kv in c.most_common(24)))
