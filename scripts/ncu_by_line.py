"""Aggregate an ncu SASS source page by CUDA source line using nvdisasm line info.
usage: ncu_by_line.py <sass_csv from `ncu --page source --csv --print-source sass`> <nvdisasm -g -c asm> <kernel substring>"""
import collections
import csv
import re
import sys

csv_path, asm_path, kname = sys.argv[1:4]
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
data = []
for r in rows[2:]:
    if not r or not r[0].startswith("0x"):
        if data:
            break
        continue
    data.append(r)
ie, it, iss = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
iw = hdr.index("L1 Wavefronts Shared")

# parse nvdisasm: find the function section for the kernel, collect (line) per instruction in order
lines = open(asm_path).read().split("\n")
in_fn, cur_line, per_inst = False, None, []
for ln in lines:
    if ln.startswith(".text.") or ln.startswith("\t.section\t.text."):
        in_fn = (kname in ln)
        continue
    if not in_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        per_inst.append(cur_line)
print("sass in csv", len(data), "sass in asm", len(per_inst))
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for r, loc in zip(data, per_inst):
    a = agg[loc]
    a[0] += int(r[ie]); a[1] += int(r[it]); a[2] += int(r[iss]); a[3] += int(r[iw])
tot = sum(a[0] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print(f"{'line':28s} {'warp inst':>12s} {'%':>6s} {'samples%':>8s} {'smem wf':>10s}")
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 60]:
    print(f"{str(loc):28s} {a[0]:12d} {100*a[0]/tot:6.1f} {100*a[2]/max(ts,1):8.1f} {a[3]:10d}")
