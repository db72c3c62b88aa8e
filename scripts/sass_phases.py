#!/usr/bin/env python
"""Static SASS census of one lk_level_kernel instantiation: instructions between barriers, by opcode.
Usage: sass_phases.py <object or .so> <WIN> <MODE> <CUMOUT 0|1>"""
import collections
import re
import subprocess
import sys

obj, win, mode, co = sys.argv[1:5]
fast = sys.argv[5] if len(sys.argv) > 5 else "0"
fun = f"_ZN3ofb15lk_level_kernelILi{win}ELi{mode}ELb{co}ELb{fast}EEEv14CUtensorMap_stS1_S1_NS_14LkKernelParamsE"
out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
ins = []
for line in out.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        ins.append(m.group(2).strip())
print(len(ins), "instructions")
seg, segs = [], []
for s in ins:
    seg.append(s)
    if "BAR.SYNC" in s:
        segs.append(seg)
        seg = []
segs.append(seg)
for k, sg in enumerate(segs):
    c = collections.Counter()
    for s in sg:
        t = s.split()
        op = t[1] if t[0].startswith("@") else t[0]
        c[op.split(".")[0]] += 1
    print(f"-- segment {k}: {len(sg)} instr: " + ", ".join(f"{a} {b}" for a, b in c.most_common(14)))
