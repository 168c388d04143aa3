"""Scan the result-invariant knobs of pmc_set_tuning (bands, prefetch, ...) and time the fused sweep.
usage: python scripts/tune_scan.py [--n N] [--phi PHI] [--delta D] [--sweeps S] [--burn B]"""
import argparse
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pmc_b200

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2 ** 20)
ap.add_argument("--phi", type=float, default=0.70)
ap.add_argument("--delta", type=float, default=0.1)
ap.add_argument("--sweeps", type=int, default=2000)
ap.add_argument("--burn", type=int, default=300)
ap.add_argument("--bands", default="1,4,6,8")
ap.add_argument("--prefetch", default="0,148,296,592")
ap.add_argument("--extra", default="")
a = ap.parse_args()

mc = pmc_b200.ParallelMC(a.n, phi=a.phi, move_delta=a.delta, n_M=4)
r = mc.rsa(seed=1234) if a.phi < 0.5 else mc.init_r()
disk, n = mc.assign(r)
mc.set_blocking(0)
mc.sweep(disk, n, 0, a.burn)
sweep = a.burn
ref_hash = None
for bands, pf in itertools.product([int(x) for x in a.bands.split(",")], [int(x) for x in a.prefetch.split(",")]):
    mc.set_tuning("bands", bands)
    mc.set_tuning("prefetch", pf)
    for name in [x for x in a.extra.split(",") if x]:
        mc.set_tuning(name, 1)
    mc.sweep(disk, n, sweep, 200)
    sweep += 200
    torch.cuda.synchronize()
    mc.reset_counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mc.sweep(disk, n, sweep, a.sweeps)
    e1.record()
    torch.cuda.synchronize()
    sweep += a.sweeps
    c = mc.counters()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"n": a.n, "phi": a.phi, "bands": bands, "prefetch": pf, "extra": a.extra,
                      "us_per_sweep": 1e3 * ms / a.sweeps, "moves_per_s": c["trials"] / (ms * 1e-3)}), flush=True)
