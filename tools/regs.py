#!/usr/bin/env python
"""Registers / spills per kernel from a `make -C svirlpool_b200/csrc` log (ptxas -v). Usage: tools/regs.py build.log [filter]"""
import re, subprocess, sys
txt = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
ents = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores.*\n.*Used (\d+) registers", txt)
dem = subprocess.run(["c++filt"] + [e[0] for e in ents], capture_output=True, text=True).stdout.split("\n")
for (n, stack, spill, regs), d in zip(ents, dem):
    d = re.sub(r"\(.*", "", d).replace("void svp::", "")
    if flt in d:
        print(f"{d:64s} regs={regs:>4s} stack={stack} spill={spill}")
