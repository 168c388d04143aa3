"""GPU parity tests (run on the B200 box): the CUDA path through the C-ABI against the CPU
oracle on the same seeded inputs.  Integer / index results and float positions are compared
BIT-EXACTLY (both sides use the same individually rounded IEEE binary32 operations)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KW = dict(phi=0.70, sigma_d=1.0, cell_w=2.0, nmax=8, n_M=4, move_delta=0.1, seed=1234)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def pair(N, **over):
    import pmc_b200
    from oracle import oracle as O
    kw = dict(KW, **over)
    return pmc_b200.ParallelMC(N, **kw), O.Oracle(N, **kw)


def assert_same_state(disk, n, odisk, on):
    assert np.array_equal(n.cpu().numpy(), on)
    d = disk.cpu().numpy()
    if not np.array_equal(bits(d), bits(odisk)):
        bad = np.nonzero((bits(d) != bits(odisk)).any(axis=(1, 2)))[0]
        raise AssertionError(f"{len(bad)} cells differ, first {bad[:8]}: gpu {d[bad[0]]} oracle {odisk[bad[0]]}")


@pytest.fixture(scope="module", autouse=True)
def _built(built):
    return built


# ---------------------------------------------------------------- init_r / assign
@pytest.mark.parametrize("N", [64, 4096, 2 ** 20])
def test_init_r_bit_exact(N):
    mc, o = pair(N)
    assert np.array_equal(bits(mc.init_r().cpu().numpy()), bits(o.init_r()))
    g = mc.geom
    assert (g.cps, g.w, g.L) == (o.g.cps, o.g.w, o.g.L)


@pytest.mark.parametrize("N,phi", [(4096, 0.70), (2 ** 16, 0.70), (2 ** 20, 0.70), (2 ** 18, 0.30)])
def test_assign_lattice_bit_exact(N, phi):
    mc, o = pair(N, phi=phi)
    disk, n = mc.assign(mc.init_r())
    odisk, on = o.assign(o.init_r())
    assert_same_state(disk, n, odisk, on)
    assert mc.counters()["status"] == 0


def test_assign_random_points_slot_order_and_lost():
    import torch
    mc, o = pair(2 ** 14, phi=0.2)
    rng = np.random.default_rng(3)
    r = ((rng.random((2, 2 ** 14)) - 0.5) * o.g.L * 1.002).astype(np.float32)   # a few outside the box
    odisk, on = o.assign(r)
    import pmc_b200
    with pytest.raises(pmc_b200.PmcError) as ei:      # blocking call + dropped particles = a return code, not silence
        mc.assign(torch.from_numpy(r).cuda())
    assert ei.value.code in (-3, -4)
    mc.strict = False                                 # the reference's own behaviour: carry on, counters tell
    mc.reset_counters()
    disk, n = mc.assign(torch.from_numpy(r).cuda())
    assert mc.last_warning in (-3, -4)
    c = mc.counters()
    assert o.lost > 0 and c["lost"] == o.lost and (c["status"] & 2)
    if not (c["status"] & 1):            # no overflow: slot order must be the reference's
        assert_same_state(disk, n, odisk, on)


def test_empty_and_ragged_cells():
    """N=16 particles in a 4x4-cell box after shifting: empty cells, ragged counts."""
    import torch
    mc, o = pair(64, phi=0.05)
    r = o.init_r()
    r[:, 40:] = r[:, :24] * 0.31 + 0.07          # pile particles up: ragged, some cells empty
    odisk, on = o.assign(r)
    disk, n = mc.assign(torch.from_numpy(r).cuda())
    assert_same_state(disk, n, odisk, on)
    assert (on == 0).any()
    mc.sweep(disk, n, 0, 7)
    o.sweep(odisk, on, 0, 7)
    assert_same_state(disk, n, odisk, on)


# ---------------------------------------------------------------- sub-sweep, one colour at a time
@pytest.mark.parametrize("N", [4096, 2 ** 16])
def test_subsweep_each_colour_bit_exact(N):
    mc, o = pair(N)
    disk, n = mc.assign(mc.init_r())
    odisk, on = o.assign(o.init_r())
    for sweep in range(3):
        for colour in (2, 0, 3, 1):
            off = o.colour_to_off(colour)
            mc.subsweep(disk, n, off, sweep)
            o.subsweep(odisk, on, off, sweep)
            assert_same_state(disk, n, odisk, on)
    c = mc.counters()
    assert (c["trials"], c["accepted"]) == (o.trials.value, o.accepted.value)


@pytest.mark.parametrize("f,dfrac", [(0, 0.25), (0, -0.31), (1, 0.5), (1, -0.4999), (0, 0.0)])
def test_shift_cells_bit_exact(f, dfrac):
    mc, o = pair(2 ** 14)
    disk, n = mc.assign(mc.init_r())
    odisk, on = o.assign(o.init_r())
    o.sweep(odisk, on, 0, 2)
    mc.sweep(disk, n, 0, 2)
    d = np.float32(dfrac * o.g.w)
    mc.shift_cells(disk, n, f, float(d))
    o.shift_cells(odisk, on, f, d)
    assert_same_state(disk, n, odisk, on)


def test_schedule_matches_oracle():
    mc, o = pair(4096)
    for s in list(range(50)) + [2 ** 32 + 5, 2 ** 40]:
        order, f, d = mc.schedule(s)
        oorder, of, od = o.schedule(s)
        assert order == oorder and f == of and np.float32(d) == od


# ---------------------------------------------------------------- the fused sweep = protocol of start.cu:237-260
@pytest.mark.parametrize("N,sweeps,over", [
    (4096, 25, {}),                                  # BASELINE config 1 geometry (cps = 32, one tile)
    (2 ** 14, 6, {}),                                # 2 x 2 tiles
    (2 ** 16, 4, {}),                                # 5 x 5 tiles, partial edge tiles
    (2 ** 16, 3, dict(n_M=7)),                       # odd n_M, more trials than particles
    (2 ** 16, 3, dict(n_M=1)),
    (2 ** 14, 4, dict(phi=0.30, move_delta=0.4)),    # dilute (config 5 regime): empty cells
    (2 ** 14, 4, dict(cell_w=1.5, phi=0.5)),         # w < 2 sigma: generic 3x3 path
    (2 ** 14, 4, dict(cell_w=2.6, phi=0.45)),        # w > 2 sigma: trials that need no neighbour column
    (2 ** 14, 3, dict(seed=2 ** 40 + 17)),
])
def test_fused_sweep_bit_exact(N, sweeps, over):
    mc, o = pair(N, **over)
    disk, n = mc.assign(mc.init_r())
    odisk, on = o.assign(o.init_r())
    mc.sweep(disk, n, 0, sweeps)
    o.sweep(odisk, on, 0, sweeps)
    assert_same_state(disk, n, odisk, on)
    c = mc.counters()
    assert (c["trials"], c["accepted"], c["lost"]) == (o.trials.value, o.accepted.value, o.lost)
    assert c["status"] == 0


@pytest.mark.parametrize("N,phi,delta,sweeps", [(2 ** 20, 0.70, 0.1, 4), (2 ** 20, 0.716, 0.1, 3), (2 ** 18, 0.30, 0.4, 4)])
def test_fused_sweep_bit_exact_at_baseline_config_sizes(N, phi, delta, sweeps):
    """BASELINE config 2 (N = 2^20, phi = 0.70: 542 x 542 cells, ~400 tiles, clipped edge tiles), the
    hexatic-region density of config 3 and the dilute regime of config 5 at sizes the oracle (OpenMP
    over same-colour cells) still finishes in seconds: positions, counts and counters bit for bit."""
    mc, o = pair(N, phi=phi, move_delta=delta)
    disk, n = mc.assign(mc.init_r())
    odisk, on = o.assign(o.init_r())
    mc.sweep(disk, n, 0, sweeps)
    o.sweep(odisk, on, 0, sweeps, omp=True)
    assert_same_state(disk, n, odisk, on)
    c = mc.counters()
    assert (c["trials"], c["accepted"], c["lost"]) == (o.trials.value, o.accepted.value, o.lost)
    assert c["status"] == 0


def test_every_colour_order_and_shift_direction_on_the_fast_path():
    """The fused kernel sizes its tiles and halos from the sweep's colour order (24 orders) and shift
    (axis, sign): 96 sweeps draw nearly all of the 96 combinations; the state must stay bit-identical
    to the oracle after every block of 16 sweeps (cps = 66: 3 x 3 tiles with clipped edge tiles)."""
    mc, o = pair(2 ** 14)
    disk, n = mc.assign(mc.init_r())
    odisk, on = o.assign(o.init_r())
    seen = set()
    for s0 in range(0, 96, 16):
        for s in range(s0, s0 + 16):
            order, f, d = o.schedule(s)
            seen.add((tuple(order), f, d > 0))
        mc.sweep(disk, n, s0, 16)
        o.sweep(odisk, on, s0, 16)
        assert_same_state(disk, n, odisk, on)
    assert len(seen) >= 55 and len({k[0] for k in seen}) == 24
    c = mc.counters()
    assert (c["trials"], c["accepted"], c["lost"]) == (o.trials.value, o.accepted.value, o.lost)


def test_fused_equals_per_call_protocol_and_split_calls():
    """pmc_sweep(K) == K x (4 x pmc_subsweep + pmc_shift_cells) == pmc_sweep(a) + pmc_sweep(K-a)."""
    mc, o = pair(2 ** 14)
    r = mc.init_r()
    d1, n1 = mc.assign(r)
    d2, n2 = mc.assign(r)
    d3, n3 = mc.assign(r)
    mc.sweep(d1, n1, 0, 5)
    for s in range(5):
        order, f, d = mc.schedule(s)
        for c in order:
            mc.subsweep(d2, n2, mc.colour_to_off(c), s)
        mc.shift_cells(d2, n2, f, d)
    mc.sweep(d3, n3, 0, 2)
    mc.sweep(d3, n3, 2, 3)
    assert_same_state(d2, n2, d1.cpu().numpy(), n1.cpu().numpy())
    assert_same_state(d3, n3, d1.cpu().numpy(), n1.cpu().numpy())


def test_garbage_in_unused_slots_is_tolerated():
    """The reference leaves garbage in slots >= n[cell] (shiftCells.h:108-110 copies it around);
    the C-ABI must not let it influence results."""
    import torch
    mc, o = pair(2 ** 14)
    disk, n = mc.assign(mc.init_r())
    odisk, on = o.assign(o.init_r())
    mask = (torch.arange(8, device="cuda")[None, :] >= n[:, None].to(torch.int64))
    junk = torch.rand_like(disk) * 2.0
    disk[:, 0, :] = torch.where(mask, junk[:, 0, :], disk[:, 0, :])
    disk[:, 1, :] = torch.where(mask, junk[:, 1, :], disk[:, 1, :])
    mc.sweep(disk, n, 0, 3)
    o.sweep(odisk, on, 0, 3)
    assert_same_state(disk, n, odisk, on)


# ---------------------------------------------------------------- observables
def test_check_and_gr_hist_bit_exact():
    mc, o = pair(2 ** 16)
    disk, n = mc.assign(mc.init_r())
    mc.sweep(disk, n, 0, 10)
    d, nn = disk.cpu().numpy(), n.cpu().numpy()
    g, oc = mc.check(disk, n), o.check(d, nn)
    assert (g["total"], g["out_of_cell"], g["overlaps"], g["bad_sentinels"]) == \
           (oc["total"], oc["out_of_cell"], oc["overlaps"], oc["bad_sentinels"])
    assert np.float32(g["min_d2"]) == oc["min_d2"]
    assert g["total"] == 2 ** 16 and g["out_of_cell"] == 0
    assert g["overlaps"] == 0 and g["min_d2"] >= 1.0     # exact on the coordinate grid (pmc.h)
    h = mc.gr_hist(disk, n, 2.0, 512)
    assert np.array_equal(h, o.gr_hist(d, nn, 2.0, 512))
    gr, gc, bp = mc.pressure_from_hist(h, 2.0, 1)
    assert gr[: 250].max() == 0.0 and 2.0 < gc < 12.0 and bp > 1.0


def test_overflow_is_reported_not_silent():
    """8 < particles in one cell: the reference writes out of bounds (start.cu:135-138);
    we must flag it."""
    import torch
    mc, o = pair(4096, sigma_d=0.01, phi=0.70 * 1e-4)
    r = o.init_r()
    r[:, :12] = r[:, :1] + (np.arange(12, dtype=np.float32) * 1e-3)[None, :]
    import pmc_b200
    with pytest.raises(pmc_b200.PmcError) as ei:
        mc.assign(torch.from_numpy(r).cuda())
    assert ei.value.code == -3                        # PMC_E_OVERFLOW
    mc.strict = False
    disk, n = mc.assign(torch.from_numpy(r).cuda())
    c = mc.counters()
    assert c["status"] & 1 and c["lost"] >= 1
    assert int(n.max()) <= 8
    # geometries whose mean occupancy leaves no head-room are refused up front (cell_w = 3 sigma at phi = 0.7: mean 8)
    with pytest.raises(pmc_b200.PmcError) as ei:
        pmc_b200.ParallelMC(4096, **dict(KW, cell_w=3.0))
    assert ei.value.code == -2


# ---------------------------------------------------------------- end to end with host buffers
def test_run_host_round_trip():
    import torch
    mc, o = pair(2 ** 14)
    r = o.init_r()
    rh = torch.from_numpy(r).pin_memory()
    g = mc.geom
    dh = torch.empty((g.local_cells, 2, 8), dtype=torch.float32).pin_memory()
    nh = torch.empty((g.local_cells,), dtype=torch.int16).pin_memory()
    mc.run_host(rh, 0, 4, dh, nh)
    odisk, on = o.assign(r)
    o.sweep(odisk, on, 0, 4)
    assert np.array_equal(nh.numpy(), on) and np.array_equal(bits(dh.numpy()), bits(odisk))
    # global coordinates back on the host (disk_to_r kernel.cu:497-507)
    disk, n = torch.from_numpy(odisk).cuda(), torch.from_numpy(on).cuda()
    rg, k = mc.disk_to_r_host(disk, n)
    ro, ko = o.disk_to_r(odisk, on)
    assert k == ko == 2 ** 14 and np.array_equal(bits(rg), bits(ro))


def test_run_host_non_blocking_on_two_handles_matches_the_oracle():
    """pmc_run_host after pmc_set_blocking(h, 0) queues H2D + assign + sweeps + D2H and returns; two handles on two
    streams (bench.py's end-to-end leg) each deliver the oracle's bits after pmc_synchronize."""
    import torch
    import pmc_b200
    from oracle import oracle as O
    kw = dict(KW, seed=77)
    N = 2 ** 16
    o = O.Oracle(N, **kw)
    r = o.init_r()
    odisk, on = o.assign(r)
    o.sweep(odisk, on, 0, 6)
    rh = torch.from_numpy(r).pin_memory()
    hs, outs, streams = [], [], [torch.cuda.Stream(), torch.cuda.Stream()]     # the handles borrow the streams: keep them alive
    for st in streams:
        with torch.cuda.stream(st):
            mc = pmc_b200.ParallelMC(N, **kw)
        mc.set_blocking(0)
        g = mc.geom
        hs.append(mc)
        outs.append((torch.empty((g.local_cells, 2, 8), dtype=torch.float32).pin_memory(),
                     torch.empty((g.local_cells,), dtype=torch.int16).pin_memory()))
    for i in range(4):                                  # two jobs per handle, back to back
        hs[i & 1].run_host(rh, 0, 6, *outs[i & 1])
    for mc in hs:
        mc.synchronize()
    for (dh, nh), mc in zip(outs, hs):
        assert np.array_equal(nh.numpy(), on) and np.array_equal(bits(dh.numpy()), bits(odisk))
        assert mc.counters()["status"] == 0 and mc.counters()["trials"] == 2 * o.trials.value


def test_disk_to_r_on_the_device_matches_the_oracle_and_the_host_variant():
    """pmc_disk_to_r (device scan + scatter) == oracle_disk_to_r (kernel.cu:497-507 order) == pmc_disk_to_r_host."""
    mc, o = pair(2 ** 16)
    disk, n = mc.assign(mc.init_r())
    mc.sweep(disk, n, 0, 5)
    r_dev, k = mc.disk_to_r(disk, n)
    r_host, kh = mc.disk_to_r_host(disk, n)
    r_or, ko = o.disk_to_r(disk.cpu().numpy(), n.cpu().numpy())
    assert k == kh == ko == 2 ** 16
    assert np.array_equal(bits(r_dev.cpu().numpy()), bits(r_or)) and np.array_equal(bits(r_host), bits(r_or))
    # (no round trip through pmc_assign: float32 GLOBAL coordinates lose the low bits of the cell-local ones)


# ---------------------------------------------------------------- the exact no-overlap invariant
@pytest.mark.parametrize("N,phi,sweeps", [(2 ** 20, 0.716, 5000), (2 ** 20, 0.70, 2000)])
def test_no_overlap_invariant_is_exact_over_thousands_of_sweeps(N, phi, sweeps):
    """north_star: "bit-exact for ... the no-overlap invariant".  On the coordinate grid (pmc.h) the
    grid shift is an exact translation and a pair's float d2 is frame independent, so after ANY
    number of sweeps pmc_check finds no pair with d2 < sigma^2 and min_d2 >= sigma^2 exactly (round 1
    drifted to 1 - 9 * 2^-24 because shiftCells re-rounded both coordinates every sweep)."""
    import pmc_b200
    mc = pmc_b200.ParallelMC(N, **dict(KW, phi=phi))
    disk, n = mc.assign(mc.init_r())
    done = 0
    for chunk in (sweeps // 5,) * 5:
        mc.sweep(disk, n, done, chunk)
        done += chunk
        g = mc.check(disk, n)
        assert g["total"] == N and g["out_of_cell"] == 0 and g["bad_sentinels"] == 0
        assert g["overlaps"] == 0 and g["min_d2"] >= 1.0, (done, g)
    c = mc.counters()
    assert c["status"] == 0 and c["lost"] == 0
    # every coordinate is still on the grid
    q = float(mc.geom.grid_q)
    x = disk.cpu().numpy()
    used = np.arange(8)[None, :] < n.cpu().numpy()[:, None]
    for dim in (0, 1):
        v = x[:, dim, :][used].astype(np.float64) / q
        assert np.array_equal(v, np.rint(v))


# ---------------------------------------------------------------- full-size properties (BASELINE sizes)
@pytest.mark.parametrize("N,phi,sweeps", [(2 ** 20, 0.70, 20), (2 ** 24, 0.70, 10), (2 ** 22, 0.30, 10)])
def test_full_size_invariants(N, phi, sweeps):
    import pmc_b200
    mc = pmc_b200.ParallelMC(N, **dict(KW, phi=phi, move_delta=0.1 if phi > 0.5 else 0.4))
    disk, n = mc.assign(mc.init_r())
    mc.sweep(disk, n, 0, sweeps)
    g = mc.check(disk, n)
    c = mc.counters()
    assert g["total"] == N and g["out_of_cell"] == 0 and g["bad_sentinels"] == 0
    # the no-overlap invariant is EXACT (coordinate grid, include/pmc.h): not one pair below sigma^2
    assert g["overlaps"] == 0 and g["min_d2"] >= 1.0 and c["status"] == 0 and c["lost"] == 0
    nonempty_bound = mc.geom.n_cells * 4 * sweeps
    assert 0 < c["trials"] <= nonempty_bound
    if phi > 0.5:
        assert c["trials"] == nonempty_bound          # no empty cells at phi = 0.70
    assert 0.05 < c["accepted"] / c["trials"] < 0.95
    # determinism: same inputs, same bits
    d2, n2 = mc.assign(mc.init_r())
    mc.sweep(d2, n2, 0, sweeps)
    import torch
    assert torch.equal(disk.view(torch.int32), d2.view(torch.int32)) and torch.equal(n, n2)


# ---------------------------------------------------------------- CUDA path vs the reference's OWN kernels
@pytest.mark.parametrize("seed", list(range(1, 17)))
def test_assign_and_shift_cells_match_the_reference_kernels_on_gpu(seed):
    """pmc_assign / pmc_shift_cells against what the reference's unmodified assign
    (kernel.cu:92-150) and V2 shiftCells (shiftCells.h:23-112) computed (tests/golden/
    ref_kernels_seed*.json, generated by oracle/ref_harness.cu on a B200)."""
    import json
    import os
    import torch
    import pmc_b200
    from test_oracle_cpu import _assert_matches_reference_step, _reference_geometry_oracle
    here = os.path.dirname(os.path.abspath(__file__))
    gold = json.load(open(os.path.join(here, "golden", f"ref_kernels_seed{seed}.json")))
    p = gold["params"]
    o = _reference_geometry_oracle(p["n_real"])
    mc = pmc_b200.ParallelMC(p["n_real"], phi=float(o.phi), sigma_d=1.0, cell_w=2.5, nmax=8, n_M=4,
                             move_delta=0.1, seed=1)
    assert mc.geom.cps == 4 and mc.geom.w == 2.5 and mc.geom.L == 10.0
    mc.strict = False       # the fixtures drop particles on purpose (lower box face, cells beyond nmax)
    r = torch.tensor(np.array(gold["r"], dtype=np.float32)[:2].copy(), device="cuda")
    disk, n = mc.assign(r)
    excess = _assert_matches_reference_step(o, disk.cpu().numpy(), n.cpu().numpy(), gold["steps"][0])
    lost0 = p["n_real"] - sum(gold["steps"][0]["n"]) + excess
    assert mc.counters()["lost"] == lost0
    for step in gold["steps"][1:]:
        if excess:
            break           # nmax 8 (ours) vs 30 (reference): states differ from the first overflow on
        mc.shift_cells(disk, n, step["f"], float(np.float32(step["d"])))
        excess = _assert_matches_reference_step(o, disk.cpu().numpy(), n.cpu().numpy(), step)
        assert mc.counters()["lost"] == lost0 + excess


# ---------------------------------------------------------------- RSA start, trajectory dump, checkpoint
def test_rsa_start_sweeps_bit_exact():
    """Dilute regime (BASELINE config 5): RSA initial configuration, empty cells, large moves."""
    mc, o = pair(2 ** 16, phi=0.30, move_delta=0.4)
    r = mc.rsa(seed=5)
    disk, n = mc.assign(r)
    odisk, on = o.assign(r.cpu().numpy())
    assert_same_state(disk, n, odisk, on)
    assert int((on == 0).sum()) > 0                  # the sparse path is exercised
    mc.sweep(disk, n, 0, 6)
    o.sweep(odisk, on, 0, 6)
    assert_same_state(disk, n, odisk, on)
    c = mc.counters()
    assert (c["trials"], c["accepted"]) == (o.trials.value, o.accepted.value)
    assert c["trials"] < 6 * 4 * o.n_cells           # empty cells perform no trials (subsweep.h:252-254)


def test_dump_matches_the_reference_format(tmp_path):
    """create_dump (kernel.cu:510-536): same header lines and row format as dumpR3.txt."""
    mc, o = pair(4096)
    disk, n = mc.assign(mc.init_r())
    path = str(tmp_path / "dump.txt")
    mc.write_dump(disk, n, path, timestep=0)
    mc.sweep(disk, n, 0, 2)
    mc.write_dump(disk, n, path, timestep=1, append=True)
    lines = open(path).read().split("\n")
    assert lines[0] == "ITEM: TIMESTEP " and lines[1] == "0" and lines[2] == "ITEM: NUMBER OF ATOMS"
    assert lines[3] == "4096" and lines[4] == "ITEM: BOX BOUNDS"
    assert lines[8] == "ITEM: ATOMS id type x y z ix iy iz"
    first = lines[9].split()
    assert first[:2] == ["1", "1"] and first[5:] == ["0", "0", "0"] and len(first) == 8
    frame = 9 + 4096
    assert lines[frame] == "ITEM: TIMESTEP " and lines[frame + 1] == "1"
    xy = np.array([[float(v) for v in ln.split()[2:4]] for ln in lines[9:9 + 4096]])
    hl = o.g.L / 2
    assert xy.min() > -hl - 1e-3 and xy.max() <= hl + 1e-3


def test_checkpoint_restart_continues_the_same_trajectory(tmp_path):
    mc, o = pair(2 ** 14)
    disk, n = mc.assign(mc.init_r())
    mc.sweep(disk, n, 0, 4)
    path = str(tmp_path / "state.ckpt")
    mc.save_checkpoint(disk, n, 4, path)
    mc.sweep(disk, n, 4, 3)
    import pmc_b200
    mc2 = pmc_b200.ParallelMC(2 ** 14, **KW)
    d2, n2, sweep = mc2.load_checkpoint(path)
    assert sweep == 4
    mc2.sweep(d2, n2, sweep, 3)
    assert_same_state(d2, n2, disk.cpu().numpy(), n.cpu().numpy())
    assert mc2.counters() == mc.counters()
    other = pmc_b200.ParallelMC(2 ** 16, **KW)
    with pytest.raises(pmc_b200.PmcError):
        other.load_checkpoint(path)                   # geometry mismatch is refused


# ---------------------------------------------------------------- observables against the hard-disk equation of state
@pytest.mark.parametrize("phi,delta", [(0.30, 0.4), (0.50, 0.2)])
def test_pressure_from_contact_value_matches_the_fluid_equation_of_state(phi, delta):
    """beta P / rho = 1 + 2 phi g(sigma+) from the g(r) histogram, against Henderson's hard-disk
    equation of state Z = (1 + phi^2 / 8) / (1 - phi)^2 (accurate to ~1 % in the fluid).
    Tolerance 3 %: statistical error of 40 samples of 65 536 disks plus the EOS's own error."""
    import pmc_b200
    N = 2 ** 16
    mc = pmc_b200.ParallelMC(N, phi=phi, sigma_d=1.0, cell_w=2.0, nmax=8, n_M=4, move_delta=delta, seed=99)
    disk, n = mc.assign(mc.rsa(seed=11) if phi < 0.5 else mc.init_r())
    mc.sweep(disk, n, 0, 3000)
    nb, rmax = 200, float(mc.geom.w) * 0.999
    hist = np.zeros(nb, dtype=np.uint64)
    sweep, samples = 3000, 40
    for _ in range(samples):
        mc.sweep(disk, n, sweep, 50)
        sweep += 50
        hist += mc.gr_hist(disk, n, rmax, nb)
    g, gc, z = mc.pressure_from_hist(hist, rmax, samples)
    z_eos = (1 + phi * phi / 8) / (1 - phi) ** 2
    chk = mc.check(disk, n)
    assert chk["total"] == N and chk["out_of_cell"] == 0
    assert abs(z - z_eos) < 0.03 * z_eos, (z, z_eos, gc)
    assert g[: int(0.95 / (rmax / nb))].sum() == 0       # no pair inside the core
    c = mc.counters()
    assert 0.2 < c["accepted"] / c["trials"] < 0.8


# ---------------------------------------------------------------- crowded cells: the 8-slot instantiation and dropped disks
def test_crowded_cells_use_the_eight_slot_path_and_match_the_oracle():
    """Small disks thrown uniformly at random: Poisson occupancy (mean 3), so cells with 7 and 8
    disks occur in most tiles (the NS = 8 instantiation of the fused sweep, mixed with NS = 6
    tiles), a few cells overflow in assign and in shiftCells (reported, identical to the oracle)."""
    import torch
    sigma, lam, N = 0.25, 3.0, 2 ** 16
    phi = lam * np.pi * sigma * sigma / 16.0          # occupancy = 16 phi / (pi sigma^2) at w = 2
    mc, o = pair(N, sigma_d=sigma, phi=float(phi), cell_w=2.0, move_delta=0.3)
    mc.strict = False                                 # overflow is part of this test: counted, not raised
    rng = np.random.default_rng(12)
    hl = np.float32(o.g.L / 2)
    r = (rng.random((2, N), dtype=np.float32) * 2 - 1) * hl * np.float32(0.9999)
    disk, n = mc.assign(torch.from_numpy(r).cuda())
    odisk, on = o.assign(r)
    assert_same_state(disk, n, odisk, on)
    assert int((on >= 7).sum()) > 50 and int(on.max()) == 8
    mc.sweep(disk, n, 0, 8)
    o.sweep(odisk, on, 0, 8)
    assert_same_state(disk, n, odisk, on)
    c = mc.counters()
    assert (c["trials"], c["accepted"], c["lost"]) == (o.trials.value, o.accepted.value, o.lost)
    assert c["lost"] > 0 and c["status"] & 1


@pytest.mark.parametrize("knob", ["force_crowded", "no_ns4", "full_halo", "generic", "four_plane", "bands1", "tile_rows8",
                                  "tile_rows28"])
def test_tuning_knobs_never_change_the_result(knob):
    """pmc_set_tuning chooses which kernel / schedule computes the sweep, never the result: "force_crowded"
    sends EVERY tile down the crowded-tile path (half-height pieces with all four planes staged), "full_halo"
    disables the per-colour-order halo planning, "generic" uses the cp.async kernel of pmc_sweep.cu, "bands"
    the stream schedule.  Bit-identical to the oracle (and hence to the default path) every time."""
    N, S = 2 ** 16, 7
    mc, o = pair(N)
    if knob == "bands1":
        mc.set_tuning("bands", 1)
    elif knob.startswith("tile_rows"):
        mc.set_tuning("tile_rows", int(knob[9:]))     # N = 2^16 runs with 16-row tiles by default (small system)
    else:
        mc.set_tuning(knob, 1)
    disk, n = mc.assign(mc.init_r())
    odisk, on = o.assign(o.init_r())
    mc.sweep(disk, n, 0, S)
    o.sweep(odisk, on, 0, S)
    assert_same_state(disk, n, odisk, on)
    c = mc.counters()
    assert (c["trials"], c["accepted"]) == (o.trials.value, o.accepted.value)
    import pmc_b200
    with pytest.raises(pmc_b200.PmcError):
        mc.set_tuning("skip_subsweeps", 1)            # nothing that changes results is reachable


def test_library_reads_no_result_changing_environment_variable():
    """Round 1 read PMC_DBG_SKIP / PMC_FAST / PMC_OVERLAP / PMC_TILE / PMC_FORCE_GENERIC from the environment in
    the product path; those strings no longer exist in the shipped library."""
    import pmc_b200
    blob = open(pmc_b200.LIB_PATH, "rb").read()
    for name in (b"PMC_DBG_SKIP", b"PMC_FAST", b"PMC_OVERLAP", b"PMC_TILE", b"PMC_FORCE_GENERIC"):
        assert name not in blob, name


def test_cells_that_become_crowded_on_the_fast_path():
    """Mildly crowded system (Poisson occupancy, mean 1.0: one cell in 10^4 holds 7 or 8 disks): most tiles hold no cell with 7 or 8
    disks and run the 3-plane fast path, but shiftCells keeps producing such cells there (their P3
    chunk goes straight to HBM and the block is flagged for the next sweep), and flagged tiles
    fall back to the 4-plane half tiles.  20 sweeps, bit-identical to the oracle throughout."""
    import torch
    sigma, lam, N = 0.25, 1.0, 2 ** 17
    phi = lam * np.pi * sigma * sigma / 16.0
    mc, o = pair(N, sigma_d=sigma, phi=float(phi), cell_w=2.0, move_delta=0.3)
    rng = np.random.default_rng(5)
    hl = np.float32(o.g.L / 2)
    r = (rng.random((2, N), dtype=np.float32) * 2 - 1) * hl * np.float32(0.9999)
    disk, n = mc.assign(torch.from_numpy(r).cuda())
    odisk, on = o.assign(r)
    assert_same_state(disk, n, odisk, on)
    seen7 = 0
    for s in range(0, 20, 4):
        mc.sweep(disk, n, s, 4)
        o.sweep(odisk, on, s, 4)
        assert_same_state(disk, n, odisk, on)
        seen7 += int((on >= 7).sum())
    assert seen7 > 10 and int((on >= 7).sum()) < on.size // 500
    c = mc.counters()
    assert (c["trials"], c["accepted"], c["lost"]) == (o.trials.value, o.accepted.value, o.lost)


# ---------------------------------------------------------------- 3-sigma agreement over INDEPENDENT seeds
def test_acceptance_and_contact_value_agree_within_three_sigma_over_independent_seeds():
    """BASELINE north_star: acceptance ratio, pressure and g(r) of the CUDA path within 3 sigma
    of the reference implementation (here: its CPU restatement) over independent seeds.  With
    EQUAL seeds the two are bit-identical (tests above); here the seeds differ, so this checks
    that nothing depends on a particular random stream.  BASELINE config 1 geometry (N = 4096,
    phi = 0.70), 8 seeds per side, 400 equilibration + 200 measured sweeps each."""
    import pmc_b200
    from oracle import oracle as O
    N, burn, meas, nb = 4096, 400, 200, 64
    kw = dict(KW)

    def stats_gpu(seed):
        mc = pmc_b200.ParallelMC(N, **dict(kw, seed=seed))
        disk, n = mc.assign(mc.init_r())
        mc.sweep(disk, n, 0, burn)
        mc.reset_counters()
        rmax = float(mc.geom.w) * 0.999
        hist = np.zeros(nb, dtype=np.uint64)
        for k in range(10):
            mc.sweep(disk, n, burn + k * (meas // 10), meas // 10)
            hist += mc.gr_hist(disk, n, rmax, nb)
        c = mc.counters()
        _, gc, z = mc.pressure_from_hist(hist, rmax, 10)
        return c["accepted"] / c["trials"], z

    def stats_cpu(seed):
        o = O.Oracle(N, **dict(kw, seed=seed))
        mc = pmc_b200.ParallelMC(N, **dict(kw, seed=seed))      # host-side maths only (pressure fit)
        disk, n = o.assign(o.init_r())
        o.sweep(disk, n, 0, burn, omp=True)
        t0, a0 = o.trials.value, o.accepted.value
        rmax = float(o.g.w) * 0.999
        hist = np.zeros(nb, dtype=np.uint64)
        for k in range(10):
            o.sweep(disk, n, burn + k * (meas // 10), meas // 10, omp=True)
            hist += o.gr_hist(disk, n, rmax, nb)
        _, gc, z = mc.pressure_from_hist(hist, rmax, 10)
        return (o.accepted.value - a0) / (o.trials.value - t0), z

    g = np.array([stats_gpu(1000 + s) for s in range(8)])
    c = np.array([stats_cpu(2000 + s) for s in range(8)])
    for col, name in ((0, "acceptance ratio"), (1, "beta P / rho")):
        dm = abs(g[:, col].mean() - c[:, col].mean())
        se = np.sqrt(g[:, col].var(ddof=1) / 8 + c[:, col].var(ddof=1) / 8)
        assert dm < 3.0 * se + 1e-12, (name, g[:, col].mean(), c[:, col].mean(), se)


# ---------------------------------------------------------------- sub-sweep vs the reference's OWN device functions
def test_subsweep_acceptance_matches_the_reference_device_functions():
    """pmc_subsweep on the states of tests/golden/trial_probes.json ("trajectory" family: every sub-sweep of
    two short runs in the reference's own geometry L = 10, w = 2.5): for every trial, the proposal is accepted
    exactly when the reference's unmodified out_of_bound / calculate_energy_in_cell /
    calculate_energy_in_neighbors (subsweep.h:73-172, run on a B200: tests/golden/ref_trials.json) say "in
    bounds and no pair energy > 0", and the resulting cells hold exactly the accepted proposals."""
    import json
    import os
    import torch
    import pmc_b200
    here = os.path.dirname(os.path.abspath(__file__))
    probes = json.load(open(os.path.join(here, "golden", "trial_probes.json")))
    ref = json.load(open(os.path.join(here, "golden", "ref_trials.json")))["probes"]
    by_episode = {}
    for p, r in zip(probes["probes"], ref):
        if p["family"] == "trajectory":
            by_episode.setdefault(p["episode"], []).append((p, r))
    assert len(by_episode) == len(probes["episodes"]) >= 20
    handles = {}
    n_trials = 0
    for e, ep in enumerate(probes["episodes"]):
        key = (ep["n_M"], ep["move_delta"], ep["seed"], ep["phi"])
        if key not in handles:
            handles[key] = pmc_b200.ParallelMC(64, phi=ep["phi"], sigma_d=1.0, cell_w=2.5, nmax=8, n_M=ep["n_M"],
                                               move_delta=ep["move_delta"], seed=ep["seed"])
            assert handles[key].geom.cps == 4 and handles[key].geom.w == 2.5
        mc = handles[key]

        def arrays(n_list, cells):
            d = np.zeros((16, 2, 8), dtype=np.float32)
            d[:, 0, :] = 1.0e18
            for c in range(16):
                for dim in (0, 1):
                    d[c, dim, :n_list[c]] = np.array(cells[c][dim], dtype=np.float32)
            return d, np.array(n_list, dtype=np.int16)
        d0, n0 = arrays(ep["n"], ep["disk"])
        disk, n = torch.tensor(d0, device="cuda"), torch.tensor(n0, device="cuda")
        mc.reset_counters()
        mc.subsweep(disk, n, ep["off"], ep["sweep"])
        c = mc.counters()
        # what the REFERENCE's functions decided for the trials of this sub-sweep
        ref_accepts = 0
        expect = d0.copy()
        for p, r in sorted(by_episode[e], key=lambda pr: (pr[0]["cy"], pr[0]["cx"], pr[0]["trial"])):
            px, py = np.float32(p["px"]), np.float32(p["py"])
            lower_face = px == 0.0 or py == 0.0            # the reference's interval is closed
            acc = (not r["oob"]) and (not r["hit"]) and not lower_face
            cell = p["cx"] + 4 * p["cy"]
            own = np.array(p["own"], dtype=np.float32)     # the cell as the trial saw it (after the shuffle)
            expect[cell, 0, :len(own)], expect[cell, 1, :len(own)] = own[:, 0], own[:, 1]
            if acc:
                expect[cell, 0, p["slot"]], expect[cell, 1, p["slot"]] = px, py
                ref_accepts += 1
            n_trials += 1
        assert c["trials"] == len(by_episode[e]) and c["accepted"] == ref_accepts == ep["accepted"], (e, c, ref_accepts)
        # the cells after the sub-sweep: last trial's view of each cell + its own outcome
        got = disk.cpu().numpy()
        last = {}
        for p, r in by_episode[e]:
            cell = p["cx"] + 4 * p["cy"]
            if cell not in last or p["trial"] > last[cell][0]["trial"]:
                last[cell] = (p, r)
        for cell, (p, r) in last.items():
            own = np.array(p["own"], dtype=np.float32)
            px, py = np.float32(p["px"]), np.float32(p["py"])
            if (not r["oob"]) and (not r["hit"]) and not (px == 0.0 or py == 0.0):
                own[p["slot"]] = (px, py)
            assert np.array_equal(bits(got[cell, 0, :len(own)]), bits(own[:, 0])), (e, cell)
            assert np.array_equal(bits(got[cell, 1, :len(own)]), bits(own[:, 1])), (e, cell)
        d1, n1 = arrays(ep["n_after"], ep["disk_after"])
        assert np.array_equal(bits(got), bits(d1)) and np.array_equal(n.cpu().numpy(), n1)
    assert n_trials == 576


# ---------------------------------------------------------------- the reference's Gaussian proposal (statistical parity)
def test_gaussian_proposal_agrees_with_the_oracle_within_three_sigma():
    """PMC_PROPOSAL_GAUSSIAN = the reference's make_move (subsweep.h:64: x + curand_normal * sigma).  It needs
    logf / sincospif, which no CPU reproduces bit for bit, so this option is checked the way north_star asks for
    the reference itself: acceptance ratio and contact value within 3 sigma over independent seeds.  Tolerance:
    |mean_gpu - mean_cpu| < 3 sqrt(se_gpu^2 + se_cpu^2), se = standard error over 8 seeds."""
    from oracle import oracle as O
    import pmc_b200
    N, burn, S = 4096, 150, 150
    kw = dict(KW, move_delta=0.08)
    acc = {"gpu": [], "cpu": []}
    gc = {"gpu": [], "cpu": []}
    for k in range(8):
        mc = pmc_b200.ParallelMC(N, **dict(kw, seed=500 + k), proposal=1)
        disk, n = mc.assign(mc.init_r())
        mc.sweep(disk, n, 0, burn)
        mc.reset_counters()
        hist = np.zeros(256, dtype=np.uint64)
        for s in range(10):
            mc.sweep(disk, n, burn + s * (S // 10), S // 10)
            hist += mc.gr_hist(disk, n, 2.0, 256)
        c = mc.counters()
        chk = mc.check(disk, n)
        assert chk["overlaps"] == 0 and chk["min_d2"] >= 1.0 and chk["total"] == N and c["status"] == 0
        acc["gpu"].append(c["accepted"] / c["trials"])
        gc["gpu"].append(mc.pressure_from_hist(hist, 2.0, 10)[1])
        o = O.Oracle(N, **dict(kw, seed=700 + k), proposal=1)       # independent seeds on the CPU side
        od, on = o.assign(o.init_r())
        o.sweep(od, on, 0, burn)
        a0, t0 = o.accepted.value, o.trials.value
        hist = np.zeros(256, dtype=np.uint64)
        for s in range(10):
            o.sweep(od, on, burn + s * (S // 10), S // 10)
            hist += o.gr_hist(od, on, 2.0, 256)
        acc["cpu"].append((o.accepted.value - a0) / (o.trials.value - t0))
        gc["cpu"].append(mc.pressure_from_hist(hist, 2.0, 10)[1])
    for name, v in (("acceptance", acc), ("g(sigma+)", gc)):
        g_, c_ = np.array(v["gpu"]), np.array(v["cpu"])
        se = np.sqrt(g_.var(ddof=1) / len(g_) + c_.var(ddof=1) / len(c_))
        assert abs(g_.mean() - c_.mean()) < 3.0 * se, (name, g_.mean(), c_.mean(), se)
    assert 0.15 < np.mean(acc["gpu"]) < 0.6
