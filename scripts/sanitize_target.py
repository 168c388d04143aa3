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
print("ok", int(h.sum()))
