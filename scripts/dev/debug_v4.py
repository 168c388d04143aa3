"""Development aid (needs a library built with -DPMC_DEBUG): isolate the phases of the v4 fused sweep with PMC_DBG_SKIP
(1 = no sub-sweeps, 2 = no shift, 4 = no store; 8, 32, 64: see pmc_internal.cuh) and compare against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import pmc_b200
from oracle import oracle as O

skip = int(os.environ.get("PMC_DBG_SKIP", "0"))  # 8 = every tile takes the crowded-tile path, 32 = no flag lookup, 64 = full halo for every colour order
N = int(os.environ.get("PMC_N", 2 ** 14))
S = int(os.environ.get("PMC_S", 3))
kw = dict(phi=0.70, sigma_d=1.0, cell_w=2.0, nmax=8, n_M=4, move_delta=0.1, seed=1234)
mc, o = pmc_b200.ParallelMC(N, **kw), O.Oracle(N, **kw)
disk, n = mc.assign(mc.init_r())
odisk, on = o.assign(o.init_r())
mc.sweep(disk, n, 0, S)
for s in range(S):
    order, f, d = o.schedule(s)
    if not (skip & 1):
        for c in order:
            o.subsweep(odisk, on, o.colour_to_off(c), s)
    if not (skip & 2):
        o.shift_cells(odisk, on, f, d)
    print("sweep", s, "order", order, "f", f, "d", d)
gn, gd = n.cpu().numpy(), disk.cpu().numpy()
cps = o.cps
bad = np.nonzero(gn != on)[0]
print("skip", skip, "cps", cps, "count mismatches", len(bad), "sum gpu", gn.sum(), "sum oracle", on.sum(), mc.counters())
if len(bad):
    for b in bad[:12]:
        print(" cell", b, "(", b % cps, b // cps, ") gpu n", gn[b], "oracle n", on[b])
    ys = np.unique(bad // cps); xs = np.unique(bad % cps)
    print(" bad rows", ys[:40], "bad cols", xs[:40])
else:
    dbad = np.nonzero((gd.view(np.uint32) != odisk.view(np.uint32)).any(axis=(1, 2)))[0]
    print("position mismatches", len(dbad))
    for b in dbad[:6]:
        print(" cell", b, "(", b % cps, b // cps, ")\n  gpu", gd[b], "\n  ora", odisk[b])
    if len(dbad):
        print(" bad rows", np.unique(dbad // cps)[:40], "bad cols", np.unique(dbad % cps)[:40])
