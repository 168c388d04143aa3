"""Aggregate an ncu SASS source page by source REGION (the device function a line belongs to).
usage: ncu_by_region.py <sass_csv of ONE kernel> <asm from nvdisasm -g -c> <kernel substring>"""
import collections, csv, os, re, sys
csv_path, asm_path, kname = sys.argv[1:4]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, "parallel-monte-carlo_b200", "csrc", "pmc_sweep4.cu")).read().split("\n")
marks = []
for i, ln in enumerate(src, 1):
    if ln.startswith("__device__") or ln.startswith("sweep4_kernel") or ln.startswith("__global__"):
        m = re.search(r"(\w+)\s*\(", ln)
        if m: marks.append((i, m.group(1)))
# finer regions inside colour_pass by comment anchors
anchors = []
for i, ln in enumerate(src, 1):
    for key, name in (("auto neighbours_min_d2", "colour:nbr-select"), ("float ox[8]", "colour:philox+shuffle"),
                      ("// trials 0..3 move slot", "colour:trials"), ("// cpy_D_sh_to_Disk", "colour:store-own")):
        if key in ln: anchors.append((i, name))
def region(loc):
    f, l = loc
    if f == "pmc_internal.cuh": return "philox" if 120 <= l <= 140 else "internal.cuh"
    if f == "sm_100_rt.hpp": return "packed fp32 (pair tests)"
    if f != "pmc_sweep4.cu": return f
    name = "top"
    for i, n in marks:
        if i <= l: name = n
    if name == "colour_pass":
        sub = "colour:setup"
        for i, n in anchors:
            if i <= l: sub = n
        return sub
    return name
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
data = [r for r in rows[2:] if r and r[0].startswith("0x")]
in_fn, cur, per = False, None, []
for ln in open(asm_path):
    if ln.startswith(".text.") or ln.startswith("\t.section\t.text."):
        in_fn = kname in ln; continue
    if not in_fn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln): per.append(cur)
assert len(per) == len(data), (len(per), len(data))
ci, cs, cw = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
st_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
for r, loc in zip(data, per):
    a = agg[region(loc)]
    a[0] += int(r[ci]); a[1] += int(r[cs]); a[2] += int(r[cw])
    for i in st_cols:
        if r[i]: a[3][hdr[i][6:]] += int(r[i])
tot, ts, tw = (sum(a[k] for a in agg.values()) for k in range(3))
print(f"total warp inst {tot}  samples {ts}  smem wavefronts {tw}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    top = ", ".join(f"{n} {100*v/ts:.1f}" for n, v in a[3].most_common(4))
    print(f"{k:28s} inst {a[0]:11d} {100*a[0]/tot:5.1f}%  samples {100*a[1]/ts:5.1f}%  smem wf {a[2]:9d} | {top}")
