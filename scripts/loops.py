"""All loops (backward branches) of a kernel in a .so with their instruction mix and fma/alu pipe estimates.
usage: loops.py lib.so <mangled substring> [min_instr]"""
import subprocess, sys, re, collections
FMA = {"FFMA", "FMUL", "FADD", "IMAD", "FFMA2", "FMUL2", "FADD2", "HFMA2", "IDP"}
ALU = {"MOV", "IADD3", "LOP3", "SHF", "PRMT", "FMNMX", "FSEL", "SEL", "ISETP", "FSETP", "PLOP3", "LEA", "IABS", "FCHK", "I2FP", "VIADD", "IADD", "FMNMX3", "IMNMX", "VIMNMX", "P2R", "R2P", "CS2R", "BMSK", "SGXT", "LOP", "FSET"}
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
minn = int(sys.argv[3]) if len(sys.argv) > 3 else 100
cur = None; funcs = {}
for l in out.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m: cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and cur: funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if sys.argv[2] not in name: continue
    print(name[:100], "total", len(ins))
    loops = []
    for a, t in ins:
        m = re.search(r"BRA\S*\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a: loops.append((tgt, a))
    for tgt, a in loops:
        body = [t for x, t in ins if tgt <= x <= a]
        if len(body) < minn: continue
        c = collections.Counter()
        for t in body:
            tk = t.split(); o = tk[1] if tk[0].startswith("@") else tk[0]
            c[o.split(".")[0]] += 1
        fma = sum(n for o, n in c.items() if o in FMA); alu = sum(n for o, n in c.items() if o in ALU)
        print(" loop %#x-%#x: %d instr | fma-pipe %d alu-pipe %d other %d" % (tgt, a, len(body), fma, alu, len(body) - fma - alu))
        print("   ", ", ".join("%s %d" % kv for kv in c.most_common(30)))
