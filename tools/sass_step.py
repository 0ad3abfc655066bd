#!/usr/bin/env python
"""Prints one steady-state step of a k_star2 variant from an object file / library (memory, FMA, barrier and branch
instructions only, in program order).  Usage: sass_step.py OBJ KERNEL_SUBSTRING [which-step] [--all]"""
import re, subprocess, sys
obj, key = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 3
elf = subprocess.run(["cuobjdump", "-elf", obj], capture_output=True, text=True).stdout
name = sorted(set(m for m in re.findall(r"\.text\.(\w+)", elf) if key in m))[0]
sass = subprocess.run(["cuobjdump", "-sass", "-fun", name, obj], capture_output=True, text=True).stdout
ins = []
for l in sass.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m: ins.append((m.group(1), m.group(2).strip()))
waits = [i for i, (_, t) in enumerate(ins) if "PHASECHK" in t]
big = [w for k, w in enumerate(waits[:-1]) if waits[k + 1] - w > 150]      # waits that open a long block = steps
lo, hi = big[which], big[which + 1]
show_all = "--all" in sys.argv
pat = re.compile(r"LDS|LDL|STL|DFMA|DADD|DMUL|FFMA|FADD|FMUL|SYNCS|STG|LDG|BRA|BSSY|SHFL")
n = {}
for a, t in ins[lo - 12:hi]:
    op = t.split()[1] if t.startswith("@") else t.split()[0]
    n[op.split(".")[0]] = n.get(op.split(".")[0], 0) + 1
    if show_all or pat.search(t): print(a, t)
print({k: v for k, v in sorted(n.items(), key=lambda x: -x[1])})
