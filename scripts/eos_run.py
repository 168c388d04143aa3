"""Equation-of-state run (BASELINE.json config 3: hard disks N = 16M, phi = 0.716, 1 B200): acceptance ratio,
g(r) and the contact-value pressure  beta P / rho = 1 + 2 phi g(sigma+)  with error bars over independent seeds.

    python scripts/eos_run.py [--n N] [--phi PHI] [--seeds K] [--burn B] [--samples S] [--stride T] [--out FILE]
                              [--oracle-n M]     # also run the CPU oracle at N = M with the first seed and
                                                 # require bit-identical histograms (small M only)

Per seed: init_r lattice -> B burn-in sweeps -> S samples, T sweeps apart, of the pair histogram (pmc_gr_hist,
r < 2 sigma, 2048 bins) -> pmc_pressure_from_hist.  Reported: mean and standard error over seeds.
Context (literature, not the reference): liquid-hexatic coexistence of hard disks at 0.700 <= phi <= 0.716 with
beta P sigma^2 = 9.185 (Bernard & Krauth, PRL 107, 155704 (2011); Engel, Anderson et al., PRE 87, 042134 (2013)),
i.e. beta P / rho = 9.185 / (4 phi / pi) = 10.08 at phi = 0.716 and 10.31 at phi = 0.700.  A run this short started
from the square lattice samples the melted, locally equilibrated state; it is a consistency check of the
observable kernels at full size, not a determination of the coexistence pressure."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pmc_b200

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2 ** 24)
ap.add_argument("--phi", type=float, default=0.716)
ap.add_argument("--delta", type=float, default=0.1)
ap.add_argument("--seeds", type=int, default=8)
ap.add_argument("--burn", type=int, default=20000)
ap.add_argument("--samples", type=int, default=20)
ap.add_argument("--stride", type=int, default=100)
ap.add_argument("--bins", type=int, default=2048)
ap.add_argument("--oracle-n", type=int, default=0)
ap.add_argument("--out", default=None)
a = ap.parse_args()
RMAX = 2.0


def run_gpu(n, seed):
    mc = pmc_b200.ParallelMC(n, phi=a.phi, move_delta=a.delta, n_M=4, seed=seed)
    disk, cnt = mc.assign(mc.init_r())
    mc.set_blocking(0)
    mc.sweep(disk, cnt, 0, a.burn)
    mc.reset_counters()
    hist = np.zeros(a.bins, dtype=np.uint64)
    sweep = a.burn
    for s in range(a.samples):
        mc.sweep(disk, cnt, sweep, a.stride)
        sweep += a.stride
        hist += mc.gr_hist(disk, cnt, RMAX, a.bins)
    c = mc.counters()
    chk = mc.check(disk, cnt)
    g, gc, bp = mc.pressure_from_hist(hist, RMAX, a.samples)
    assert c["status"] == 0 and chk["overlaps"] == 0 and chk["total"] == n and chk["min_d2"] >= 1.0
    mc.close()
    return dict(seed=seed, acceptance=c["accepted"] / c["trials"], g_contact=gc, beta_p_over_rho=bp,
                g=g, hist=hist, min_d2=chk["min_d2"])


def summarise(rows, key):
    v = np.array([r[key] for r in rows], dtype=np.float64)
    return float(v.mean()), float(v.std(ddof=1) / np.sqrt(len(v))) if len(v) > 1 else float("nan")


t0 = time.time()
rows = [run_gpu(a.n, 1000 + k) for k in range(a.seeds)]
torch.cuda.synchronize()
wall = time.time() - t0
out = dict(config=dict(n_particles=a.n, phi=a.phi, move_delta=a.delta, n_M=4, cell_w=2.0, seeds=[r["seed"] for r in rows],
                       burn_in_sweeps=a.burn, samples=a.samples, stride_sweeps=a.stride, r_max=RMAX, bins=a.bins,
                       init="square lattice (init_r)"),
           wall_s=wall)
for key in ("acceptance", "g_contact", "beta_p_over_rho"):
    m, e = summarise(rows, key)
    out[key] = dict(mean=m, stderr=e, per_seed=[r[key] for r in rows])
out["beta_p_sigma2"] = dict(mean=out["beta_p_over_rho"]["mean"] * 4 * a.phi / np.pi,
                            stderr=out["beta_p_over_rho"]["stderr"] * 4 * a.phi / np.pi)
G = np.array([r["g"] for r in rows])
dr = RMAX / a.bins
keep = slice(int(0.98 / dr), a.bins, 8)
out["g_of_r"] = dict(r=[float((i + 0.5) * dr) for i in range(a.bins)][keep],
                     mean=[float(x) for x in G.mean(axis=0)[keep]],
                     stderr=[float(x) for x in (G.std(axis=0, ddof=1) / np.sqrt(len(rows)))[keep]] if len(rows) > 1 else None)
out["min_d2"] = min(r["min_d2"] for r in rows)
out["literature"] = dict(beta_p_sigma2_coexistence=9.185, phi_liquid=0.700, phi_hexatic=0.716,
                         beta_p_over_rho_at_phi=9.185 / (4 * a.phi / np.pi),
                         source="Bernard & Krauth 2011; Engel, Anderson et al. 2013 (context only, not the reference)")
if a.oracle_n:
    from oracle import oracle as O
    n = a.oracle_n
    g1 = run_gpu(n, 1000)
    o = O.Oracle(n, phi=a.phi, move_delta=a.delta, n_M=4, seed=1000)
    d, c = o.assign(o.init_r())
    o.sweep(d, c, 0, a.burn, omp=True)
    acc0, tr0 = o.accepted.value, o.trials.value
    hist = np.zeros(a.bins, dtype=np.uint64)
    sweep = a.burn
    for s in range(a.samples):
        o.sweep(d, c, sweep, a.stride, omp=True)
        sweep += a.stride
        hist += o.gr_hist(d, c, RMAX, a.bins)
    same = bool(np.array_equal(hist, g1["hist"]))
    out["oracle_check"] = dict(n_particles=n, seed=1000, histograms_bit_identical=same,
                               acceptance_oracle=(o.accepted.value - acc0) / (o.trials.value - tr0),
                               acceptance_gpu=g1["acceptance"], beta_p_over_rho=g1["beta_p_over_rho"])
    assert same, "CUDA path and oracle disagree"
text = json.dumps(out)
if a.out:
    open(a.out, "w").write(text + "\n")
print(json.dumps({k: v for k, v in out.items() if k != "g_of_r"}))
