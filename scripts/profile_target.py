"""Short, deterministic target for ncu: N=2^24 phi=0.70, a few fused sweeps after a short burn-in."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pmc_b200

N = int(os.environ.get("PMC_N", 2 ** 24))
BURN = int(os.environ.get("PMC_BURN", 20))
SWEEPS = int(os.environ.get("PMC_SWEEPS", 5))
mc = pmc_b200.ParallelMC(N, phi=0.70, move_delta=0.1, n_M=4)
disk, n = mc.assign(mc.init_r())
mc.sweep(disk, n, 0, BURN)
mc.sweep(disk, n, BURN, SWEEPS)
torch.cuda.synchronize()
c = mc.counters()
print("ok", c)
assert c["status"] == 0
