"""Aggregate an ncu SASS source page by CUDA source line using nvdisasm line info.
usage: ncu_by_line.py <sass_csv from `ncu -i rep --page source --csv --print-source sass`>
                      <asm from `nvdisasm -g -c cubin`> <kernel substring> [top N]
The cubin must be the build that was profiled (instruction counts are checked)."""
import collections
import csv
import re
import sys

csv_path, asm_path, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
data = []
for r in rows[2:]:
    if not r or not r[0].startswith("0x"):
        if data:
            break
        continue
    data.append(r)
col = {k: hdr.index(k) for k in ("Instructions Executed", "Thread Instructions Executed", "# Samples",
                                 "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "Source")}

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
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(\S+)", ln)
    if m:
        per_inst.append((cur_line, m.group(1)))
print("sass in csv", len(data), "sass in asm", len(per_inst))
if len(data) != len(per_inst):
    print("WARNING: instruction counts differ: the cubin is not the profiled build")
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
ops = collections.defaultdict(int)
for r, (loc, op) in zip(data, per_inst):
    a = agg[loc]
    ie = int(r[col["Instructions Executed"]])
    a[0] += ie
    a[1] += int(r[col["Thread Instructions Executed"]])
    a[2] += int(r[col["# Samples"]])
    a[3] += int(r[col["L1 Wavefronts Shared"]])
    a[4] += int(r[col["L1 Wavefronts Shared Ideal"]])
    ops[op.split(".")[0].rstrip(";")] += ie
tot = sum(a[0] for a in agg.values())
ts = sum(a[2] for a in agg.values())
tw = sum(a[3] for a in agg.values())
print(f"total warp inst {tot}  samples {ts}  smem wavefronts {tw} (ideal {sum(a[4] for a in agg.values())})")
print(f"{'line':28s} {'warp inst':>12s} {'%':>6s} {'samples%':>8s} {'smem wf':>10s} {'ideal':>10s}")
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{str(loc):28s} {a[0]:12d} {100*a[0]/tot:6.1f} {100*a[2]/max(ts,1):8.1f} {a[3]:10d} {a[4]:10d}")
print("\nby opcode (warp instructions):")
for op, n in sorted(ops.items(), key=lambda kv: -kv[1])[:30]:
    print(f"  {op:12s} {n:12d} {100*n/tot:6.1f}%")
