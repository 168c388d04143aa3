"""3-D Lennard-Jones mode, GPU side (include/pmc_lj.h): the CUDA path through the C-ABI against the CPU oracle,
BIT FOR BIT on the default (uniform-cube) proposal, and against what the reference's own 3-D kernels computed."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CONFIGS = [dict(n_particles=64, L=10.0, cells_per_side=4, nmax=10, n_M=10, sigma=0.5, beta=0.3),      # start.cu:14-24
           dict(n_particles=1000, L=10.0, cells_per_side=4, nmax=30, n_M=15, sigma=0.5, beta=0.3),   # the run of slide 14
           dict(n_particles=8000, L=24.0, cells_per_side=8, nmax=32, n_M=12, sigma=0.3, beta=0.8)]


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module", autouse=True)
def _built(built):
    return built


def pair(cfg, **over):
    import pmc_b200
    from oracle import oracle as O
    kw = dict(cfg, **over)
    n = kw.pop("n_particles")
    return pmc_b200.ParallelMCLJ(n, **kw), O.OracleLJ(n, **kw)


def same_state(disk, n, odisk, on):
    assert np.array_equal(n.cpu().numpy(), on)
    d = disk.cpu().numpy()
    used = np.arange(d.shape[2])[None, None, :] < on[:, None, None]
    assert np.array_equal(bits(d)[np.broadcast_to(used, d.shape)], bits(odisk)[np.broadcast_to(used, d.shape)])


@pytest.mark.parametrize("cfg", CONFIGS)
def test_lj_init_assign_subsweep_shift_bit_exact(cfg):
    mc, o = pair(cfg)
    r = mc.init_r()
    assert np.array_equal(bits(r.cpu().numpy()), bits(o.init_r()))
    disk, n = mc.assign(r)
    odisk, on = o.assign(o.init_r())
    same_state(disk, n, odisk, on)
    for sweep in range(3):
        order, f, d = mc.schedule(sweep)
        assert (order, f, np.float32(d)) == o.schedule(sweep)
        for colour in order:
            off = mc.colour_to_off(colour)
            assert off == o.colour_to_off(colour)
            mc.subsweep(disk, n, off, sweep)
            o.subsweep(odisk, on, off, sweep)
            same_state(disk, n, odisk, on)
        mc.shift_cells(disk, n, f, d)
        o.shift_cells(odisk, on, f, d)
        same_state(disk, n, odisk, on)
    c = mc.counters()
    assert (c["trials"], c["accepted"], c["lost"], c["status"]) == (o.trials.value, o.accepted.value, o.lost, 0)
    assert abs(c["dE"] - o.dE.value) <= 1e-9 * max(1.0, abs(o.dE.value))


@pytest.mark.parametrize("cfg,sweeps", [(CONFIGS[0], 60), (CONFIGS[1], 25), (CONFIGS[2], 10)])
def test_lj_sweep_with_energy_trace_bit_exact(cfg, sweeps):
    """pmc_lj_sweep = start.cu:237-260 with the V2 energy trace (kernel.cu:672-680): same states as the oracle, the
    per-sweep accepted energy changes equal to 1e-9 (double sums in a different order), and lattice energy + trace
    == calc_energy (kernel.cu:452-470) of the final state computed by the device kernel AND by the O(N^2) oracle."""
    mc, o = pair(cfg)
    disk, n = mc.assign(mc.init_r())
    odisk, on = o.assign(o.init_r())
    e0 = mc.energy(disk, n)
    assert abs(e0 - o.energy(odisk, on)) <= 1e-6 * abs(e0)
    tr = mc.sweep(disk, n, 0, sweeps, trace=True)
    otr = o.sweep(odisk, on, 0, sweeps)
    same_state(disk, n, odisk, on)
    assert np.allclose(tr, otr, rtol=1e-9, atol=1e-9)
    e1 = mc.energy(disk, n)
    assert abs(e1 - o.energy(odisk, on)) <= 1e-6 * abs(e1)
    assert abs((e0 + tr.sum()) - e1) <= 2e-3 * max(1.0, abs(e1))
    c = mc.counters()
    assert (c["trials"], c["accepted"]) == (o.trials.value, o.accepted.value) and c["status"] == 0
    r, k = mc.disk_to_r_host(disk, n)
    assert k == cfg["n_particles"] and np.all(np.abs(r) <= cfg["L"] / 2)


@pytest.mark.parametrize("seed", [1, 7, 11, 16])
def test_lj_assign_and_shift_cells_equal_the_reference_kernels_on_gpu(seed):
    """pmc_lj_assign / pmc_lj_shift_cells against the outputs of the reference's unmodified 3-D assign
    (kernel.cu:92-150) and V2 shiftCells (shiftCells.h:23-112): same layout, same global coordinates, bit for bit."""
    import torch
    import pmc_b200
    gold = json.load(open(os.path.join(HERE, "golden", f"ref_kernels_seed{seed}.json")))
    p = gold["params"]
    mc = pmc_b200.ParallelMCLJ(p["n_real"], L=10.0, cells_per_side=4, nmax=30)
    mc.strict = False                       # the fixtures put particles on the lower box face on purpose
    disk, n = mc.assign(torch.tensor(np.array(gold["r"], dtype=np.float32), device="cuda"))

    def same(step):
        nn, d = n.cpu().numpy(), disk.cpu().numpy()
        np.testing.assert_array_equal(nn.astype(np.int64), np.array(step["n"], dtype=np.int64))
        for c in range(64):
            if nn[c]:
                assert np.array_equal(bits(d[c, :, :nn[c]]), bits(np.array(step["cells"][str(c)], dtype=np.float32))), c
    same(gold["steps"][0])
    assert mc.counters()["lost"] == p["n_real"] - sum(gold["steps"][0]["n"])
    for step in gold["steps"][1:]:
        mc.shift_cells(disk, n, step["f"], float(np.float32(step["d"])))
        same(step)


def test_lj_gaussian_proposal_agrees_with_the_oracle_within_three_sigma():
    """PMC_PROPOSAL_GAUSSIAN = make_move of the reference (subsweep.h:64).  logf / sincospif differ between a CPU
    and the GPU, so: acceptance ratio and energy per particle within 3 sigma over 8 independent seeds per side."""
    cfg = CONFIGS[1]
    acc, en = {"gpu": [], "cpu": []}, {"gpu": [], "cpu": []}
    for k in range(8):
        mc, _ = pair(cfg, seed=100 + k, proposal=1)
        disk, n = mc.assign(mc.init_r())
        mc.sweep(disk, n, 0, 150)
        mc.reset_counters()
        es = []
        for b in range(5):
            mc.sweep(disk, n, 150 + 10 * b, 10)
            es.append(mc.energy(disk, n) / cfg["n_particles"])
        c = mc.counters()
        assert c["status"] == 0
        acc["gpu"].append(c["accepted"] / c["trials"]); en["gpu"].append(np.mean(es))
        _, o = pair(cfg, seed=300 + k, proposal=1)
        od, on = o.assign(o.init_r())
        o.sweep(od, on, 0, 150)
        a0, t0 = o.accepted.value, o.trials.value
        es = []
        for b in range(5):
            o.sweep(od, on, 150 + 10 * b, 10)
            es.append(o.energy(od, on) / cfg["n_particles"])
        acc["cpu"].append((o.accepted.value - a0) / (o.trials.value - t0)); en["cpu"].append(np.mean(es))
    for name, v in (("acceptance", acc), ("energy per particle", en)):
        g_, c_ = np.array(v["gpu"]), np.array(v["cpu"])
        se = np.sqrt(g_.var(ddof=1) / len(g_) + c_.var(ddof=1) / len(c_))
        assert abs(g_.mean() - c_.mean()) < 3.0 * se, (name, g_.mean(), c_.mean(), se)
