#!/usr/bin/env python
"""Registers / spills / shared memory per kernel from a `-Xptxas -v` log (build/obj/*.ptxas.log).
    python scripts/ptxas_summary.py build/obj/spmv.ptxas.log [filter]"""
import re
import subprocess
import sys

log = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
ents = re.findall(r"Compiling entry function '(\S+)' for '\S+'\n.*?Function properties for \S+\n\s*(.*?)\n.*?Used (\d+) registers(.*?)\n", log, re.S)
names = subprocess.run(["c++filt"], input="\n".join(e[0] for e in ents), capture_output=True, text=True).stdout.split("\n")
for (mangled, props, regs, rest), name in zip(ents, names):
    name = re.sub(r"b200::\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    if flt and flt not in name:
        continue
    spill = re.search(r"(\d+) bytes spill stores", props)
    smem = re.search(r"(\d+) bytes smem", rest)
    print(f"{name:70s} regs {regs:>3s} spill {spill.group(1) if spill else '?':>4s} smem {smem.group(1) if smem else '0'}")
