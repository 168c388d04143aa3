"""Short, deterministic target for ncu.

    python scripts/profile_target.py [--n N] [--phi PHI] [--delta D] [--burn B] [--sweeps S] [--bands K] [--all]

Runs B burn-in sweeps (one pmc_sweep call), then S sweeps (a second call) of the fused kernel.  --bands 1
makes one launch = one whole sweep (what roofline.traffic is quoted per).  --all additionally launches every
other kernel of the library once (assign, stand-alone shiftCells, single-colour sub-sweep, check, g(r)
histogram, disk -> r), so that one launch list / one `--set full` capture covers them all.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pmc_b200

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2 ** 24)
ap.add_argument("--phi", type=float, default=0.70)
ap.add_argument("--delta", type=float, default=0.1)
ap.add_argument("--burn", type=int, default=20)
ap.add_argument("--sweeps", type=int, default=5)
ap.add_argument("--bands", type=int, default=0)
ap.add_argument("--all", action="store_true")
a = ap.parse_args()

mc = pmc_b200.ParallelMC(a.n, phi=a.phi, move_delta=a.delta, n_M=4)
if a.bands:
    mc.set_tuning("bands", a.bands)
r = mc.rsa(seed=1234) if a.phi < 0.5 else mc.init_r()
disk, n = mc.assign(r)
mc.sweep(disk, n, 0, a.burn)
mc.sweep(disk, n, a.burn, a.sweeps)
torch.cuda.synchronize()
c = mc.counters()
print("ok", c, "acceptance", c["accepted"] / max(c["trials"], 1))
assert c["status"] == 0
if a.all:
    order, f, d = mc.schedule(12345)
    mc.subsweep(disk, n, mc.colour_to_off(order[0]), 12345)
    mc.shift_cells(disk, n, f, d)
    chk = mc.check(disk, n)
    h = mc.gr_hist(disk, n, 2.0, 1024)
    r2, k = mc.disk_to_r(disk, n)
    d2, n2 = mc.assign(r)
    torch.cuda.synchronize()
    print("all ok", chk, int(h.sum()), k)
    assert chk["overlaps"] == 0 and k == a.n
