import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pmc_b200
from oracle import oracle as O
sigma, lam, N = 0.25, 3.0, 2 ** 16
phi = lam * np.pi * sigma * sigma / 16.0
kw = dict(phi=float(phi), sigma_d=sigma, cell_w=2.0, nmax=8, n_M=4, move_delta=0.3, seed=1234)
mc, o = pmc_b200.ParallelMC(N, **kw), O.Oracle(N, **kw)
rng = np.random.default_rng(12)
hl = np.float32(o.g.L / 2)
r = (rng.random((2, N), dtype=np.float32) * 2 - 1) * hl * np.float32(0.9999)
disk, n = mc.assign(torch.from_numpy(r).cuda())
odisk, on = o.assign(r)
def cmp(tag):
    gn, gd = n.cpu().numpy(), disk.cpu().numpy()
    bad = np.nonzero(gn != on)[0]
    dbad = np.nonzero((gd.view(np.uint32) != odisk.view(np.uint32)).any(axis=(1, 2)))[0]
    print(tag, "cps", o.cps, "count mismatches", len(bad), "position mismatches", len(dbad), mc.counters(), o.lost)
    for b in dbad[:4]:
        print(" cell", b, "(", b % o.cps, b // o.cps, ") n gpu/oracle", gn[b], on[b], "\n  gpu", gd[b], "\n  ora", odisk[b])
    return len(bad) + len(dbad)
cmp("assign")
S = int(os.environ.get("PMC_S", 8))
for s in range(S):
    mc.sweep(disk, n, s, 1)
    o.sweep(odisk, on, s, 1)
    if cmp("sweep %d (f,d)=%s" % (s, o.schedule(s)[1:])): break
