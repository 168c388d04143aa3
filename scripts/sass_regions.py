"""Development aid: map the SASS of one kernel of libpmc_b200.so to source regions (by -lineinfo)
and count instructions and local-memory (spill) accesses per contiguous region.
usage: sass_regions.py <kernel substring> [start:end to dump]"""
import os, re, subprocess, sys, tempfile
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "parallel-monte-carlo_b200", "libpmc_b200.so")
kname = sys.argv[1]
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "pmc_sweep4", lib], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.startswith("pmc_sweep4") and f.endswith(".cubin")][0]
asm = subprocess.check_output(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], text=True)
src = open(os.path.join(root, "parallel-monte-carlo_b200", "csrc", "pmc_sweep4.cu")).read().split("\n")
# function start lines -> names
marks = []
for i, ln in enumerate(src, 1):
    m = re.match(r"(?:__device__|__global__|static|template).*?\b(\w+)\s*\(", ln)
    if ln.startswith("__device__") or ln.startswith("sweep4_kernel") or ln.startswith("__global__"):
        m = re.search(r"(\w+)\s*\(", ln)
        if m: marks.append((i, m.group(1)))
def region(loc):
    f, l = loc
    if f != "pmc_sweep4.cu": return f
    name = "top"
    for i, n in marks:
        if i <= l: name = n
    return name
in_fn, cur, seq = False, None, []
for ln in asm.split("\n"):
    if ln.startswith(".text.") or ln.startswith("\t.section\t.text."):
        in_fn = kname in ln; continue
    if not in_fn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
    if m: seq.append((cur, m.group(2)))
print("instructions", len(seq))
if len(sys.argv) > 2:
    a, b = map(int, sys.argv[2].split(":"))
    for i in range(a, b): print(i, seq[i][0][1], seq[i][1][:100])
    sys.exit(0)
runs = []
for loc, ins in seq:
    r = region(loc)
    loc_mem = 1 if re.search(r"\b(STL|LDL)\b", ins) else 0
    if runs and (runs[-1][0] == r): runs[-1][1] += 1; runs[-1][2] += loc_mem
    else: runs.append([r, 1, loc_mem])
out = []
for r, c, l in runs:
    if out and c < 12: out[-1][1] += c; out[-1][2] += l
    elif out and out[-1][0] == r: out[-1][1] += c; out[-1][2] += l
    else: out.append([r, c, l])
pos = 0
for r, c, l in out:
    print(f"{pos:6d} {r:28s} {c:5d}  local {l}")
    pos += c
