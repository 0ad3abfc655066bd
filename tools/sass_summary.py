#!/usr/bin/env python
"""SASS summary of the shipped library (cuobjdump -sass; runs without a GPU): per kernel of interest the instruction total and
the mnemonics that identify the Blackwell path.  Usage: sass_summary.py [LIB] > profiles/<name>.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diffeqoperators.jl_b200", "libdeo_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, counts = None, {}
KEYS = ["UTMALDG", "SYNCS", "USETMAXREG", "LDGSTS", "DFMA", "FFMA2", "FFMA", "LDS", "STG", "LDL", "STL"]
for line in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        kern = m.group(1); counts[kern] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1); counts[kern]["total"] += 1
        for k in KEYS:
            if op == k or (k in ("UTMALDG", "SYNCS", "USETMAXREG", "LDGSTS") and op.startswith(k)):
                counts[kern][k] += 1
def demangle(n):
    full = subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip().replace("void deo::", "")
    return re.sub(r"\((int|bool)\)", "", full.split(">(")[0] + ">")
print(f"# SASS summary of {os.path.basename(lib)} (cuobjdump -sass, sm_100a)")
print("#   UTMALDG = TMA (cp.async.bulk.tensor), SYNCS = mbarrier ops, USETMAXREG = setmaxnreg register split, LDGSTS = cp.async (element loader),")
print("#   DFMA / FFMA / FFMA2 = the stencil arithmetic (FFMA2 = packed fma.rn.f32x2), LDL / STL = register spills")
print(f"# kernels in the library: {len(counts)}")
tot = collections.Counter()
for c in counts.values(): tot.update(c)
print("library totals: " + ", ".join(f"{k} {tot[k]}" for k in ["UTMALDG", "SYNCS", "USETMAXREG", "LDGSTS", "FFMA2"]) + "\n")
want = [("C5 / smoke", "k_star2IdLi2ELb1ELi7ELb0"), ("C3 F64", "k_star2IdLi3ELb1ELi7ELb0"), ("C3 F32", "k_star2IfLi3ELb1ELi7ELb0"), ("C2 (2-D strip)", "k_star2IdLi2ELb0ELi5ELb0"),
        ("C4 (TABLE)", "k_star2IdLi2ELb1ELi7ELb1"), ("C1", "k_lineIdLi1ELb0ELi1"), ("first generation, C5", "k_starIdLi2ELi2ELi8ELb1ELi7ELb0")]
for label, key in want:
    names = [n for n in counts if key in n]
    if not names: continue
    n = names[0]; c = counts[n]
    print(f"{label:24s} {demangle(n):40s} total {c['total']:6d}  " + "  ".join(f"{k} {c[k]}" for k in KEYS))
