"""Small deterministic target for compute-sanitizer (memcheck / racecheck): every kernel of the
path once or twice on a tiny system, both fused-sweep instantiations, both shift axes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pmc_b200
mc = pmc_b200.ParallelMC(2 ** 14, phi=0.70, move_delta=0.1, n_M=4)
disk, n = mc.assign(mc.init_r())
mc.sweep(disk, n, 0, 6)                                  # fast path: import, 6 fused sweeps (f = 0 and 1 occur), export
order, f, d = mc.schedule(6)
for c in order:
    mc.subsweep(disk, n, mc.colour_to_off(c), 6)         # generic single-colour kernel
mc.shift_cells(disk, n, f, d)
print(mc.check(disk, n), mc.counters())
h = mc.gr_hist(disk, n, 1.9, 64)
torch.cuda.synchronize()
# crowded cells: tiles that take the 4-plane half-height path, cells that become crowded on the fast path
import numpy as np
sigma, lam, N = 0.25, 1.5, 2 ** 14
mc2 = pmc_b200.ParallelMC(N, sigma_d=sigma, phi=float(lam * np.pi * sigma * sigma / 16.0), cell_w=2.0, move_delta=0.3, n_M=4)
mc2.strict = False          # overflow may happen here: counted, not raised
rng = np.random.default_rng(12)
hl = np.float32(mc2.geom.L / 2)
r = (rng.random((2, N), dtype=np.float32) * 2 - 1) * hl * np.float32(0.9999)
d2, n2 = mc2.assign(torch.from_numpy(r).cuda())
mc2.sweep(d2, n2, 0, 6)
torch.cuda.synchronize()
print("crowded", int((n2 >= 7).sum()), mc2.counters())
print("ok", int(h.sum()))
