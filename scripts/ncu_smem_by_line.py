"""Shared-memory wavefronts per (source line, opcode) from an ncu SASS source page.
usage: ncu_smem_by_line.py <sass_csv> <nvdisasm -g -c asm> <kernel substring> [top N]"""
import collections, csv, re, sys
csv_path, asm_path, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
ci = {k: hdr.index(k) for k in ["Source", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal"]}
data = []
for r in rows[2:]:
    if not r or not r[0].startswith("0x"):
        if data:
            break
        continue
    data.append(r)
in_fn, cur, per = False, None, []
for ln in open(asm_path).read().split("\n"):
    if ln.startswith(".text.") or ln.startswith("\t.section\t.text."):
        in_fn = kname in ln
        continue
    if not in_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(\S+)", ln):
        per.append(cur)
assert len(per) == len(data), (len(per), len(data))
agg = collections.defaultdict(lambda: [0, 0, 0])
for r, loc in zip(data, per):
    w = int(r[ci["L1 Wavefronts Shared"]])
    if w == 0:
        continue
    s = r[ci["Source"]].split()
    op = s[1] if s[0].startswith("@") else s[0]
    a = agg[(loc, op)]
    a[0] += int(r[ci["Instructions Executed"]]); a[1] += w; a[2] += int(r[ci["L1 Wavefronts Shared Ideal"]])
tot = sum(v[1] for v in agg.values())
print("total smem wavefronts", tot, "ideal", sum(v[2] for v in agg.values()))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{str(k):50s} inst {v[0]:9d} wf {v[1]:9d} ideal {v[2]:9d}  {100*v[1]/tot:5.1f}%")
