"""Opcode histogram of every kernel in libpmc_b200.so (cuobjdump -sass) + the lines that prove the TMA /
mbarrier / packed-FP32 path.  Usage: python scripts/sass_opcodes.py > profiles/r2/sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "parallel-monte-carlo_b200", "libpmc_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda s: subprocess.run(["cu++filt", s], capture_output=True, text=True).stdout.strip() or s
kern, hist, proof = None, collections.OrderedDict(), collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = demangle(m.group(1))
        kern = kern.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
        kern = kern[:kern.rindex(">(") + 1] if ">(" in kern else kern.split("(")[0]
        hist[kern] = collections.Counter()
        proof[kern] = []
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        hist[kern][op.split(".")[0]] += 1
        if re.match(r"UTMALDG|UTMAPF|UTMASTG|UBLKCP|SYNCS|FFMA2|FADD2|FMUL2|FMNMX3|REDUX|LDGSTS", op) and len(proof[kern]) < 400:
            proof[kern].append(op)
print("# cuobjdump -sass", os.path.relpath(lib, ROOT), "(arch sm_100a)")
print("# per kernel: total instructions, then opcodes by count; '*' lines = Blackwell-specific evidence\n")
for k, h in hist.items():
    tot = sum(h.values())
    print(f"== {k}   [{tot} instructions]")
    print("   " + "  ".join(f"{op}:{n}" for op, n in h.most_common(28)))
    ev = collections.Counter(proof[k])
    if ev:
        print("   * " + "  ".join(f"{op} x{n}" for op, n in sorted(ev.items())))
    print()
