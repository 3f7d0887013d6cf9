"""Registers / spills / stack per kernel from a `nvcc -Xptxas -v` log (development tool).
usage: python tools/ptxas_table.py <log> [name-filter]"""
import re
import subprocess
import sys

txt = open(sys.argv[1]).read()
filt = sys.argv[2] if len(sys.argv) > 2 else ""
pat = re.compile(r"Compiling entry function '(\S+)'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", re.S)
names = [m.group(1) for m in pat.finditer(txt)]
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.strip().split("\n") if names else []
for m, dn in zip(pat.finditer(txt), dem):
    dn = dn.replace("athk::", "").replace("(int)", "").replace("(bool)", "")
    if filt and filt not in dn:
        continue
    print("%-90s regs %3s stack %4s spill st/ld %4s/%4s" % (dn[:90], m.group(5), m.group(2), m.group(3), m.group(4)))
