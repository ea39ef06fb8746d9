"""Registers / spills / stack per kernel from a ptxas -v log.  usage: ptxas_summary.py log [substring]"""
import re, sys, subprocess
txt = open(sys.argv[1]).read()
pat = sys.argv[2] if len(sys.argv) > 2 else ""
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt):
    name = m.group(1)
    try:
        name = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    except Exception:
        pass
    name = name.replace("md2::", "").replace("(Params)", "")
    if pat in name:
        print("%4s regs %4s stack %4s/%4s spill  %s" % (m.group(5), m.group(2), m.group(3), m.group(4), name[:110]))
