#!/usr/bin/env python
"""Generates the (state, proposal) probes that oracle/ref_harness_v1.cu feeds to the REFERENCE's own
sub-sweep device functions (subsweep.h:73-172), and the bookkeeping the tests need to compare the
answers with the oracle (CPU) and with the CUDA path (GPU).

    python tests/golden/make_trial_probes.py          # in the build container (needs only the oracle)
        -> tests/golden/trial_probes.json              probes in OUR conventions (cell-local, 2-D)
        -> tests/golden/trial_probes_in.txt            the same in the reference's (global, 3-D) for the harness
    gpurun -- './oracle/_ref/ref_harness_v1 < tests/golden/trial_probes_in.txt > gpurun_out/ref_trials.json'
    cp gpurun_out/ref_trials.json tests/golden/ref_trials.json

Geometry = the reference's own (#define block start.cu:14-18: L = 10, w = 2.5, cellsPerSide = 4); all
coordinates are multiples of 2^-21, so local <-> global conversion is exact in binary32.

Family "dyadic": hand-built states on a 2^-10 grid, one target disk per probe in the own cell or in one of
  the 8 neighbour cells (incl. across the periodic box edge), at r = 1 exactly, 1 -+ 2^-10, well inside,
  in the attractive LJ range, plus proposals exactly on / just outside the cell faces.
Family "trajectory": every trial the oracle executes in short runs from the lattice at phi = 0.50
  (64 disks), with the state of the sub-sweep it belongs to: realistic dense neighbourhoods.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

W, LBOX, CPS = 2.5, 10.0, 4


def ref_geometry_oracle(n, **kw):
    base = np.float32(n * np.pi / 400.0)
    for cand in (base, np.nextafter(base, np.float32(0)), np.nextafter(base, np.float32(1))):
        try:
            o = O.Oracle(n, phi=float(cand), sigma_d=1.0, cell_w=2.5, nmax=8, **kw)
        except ValueError:
            continue
        if o.cps == 4 and o.g.w == 2.5 and o.g.L == 10.0:
            return o, float(cand)
    raise AssertionError("no float32 phi reproduces L=10, w=2.5")


def to_global(cell, x, y):
    cx, cy = cell % CPS, cell // CPS
    gx, gy = cx * W - LBOX / 2 + float(x), cy * W - LBOX / 2 + float(y)
    assert float(np.float32(gx)) == gx and float(np.float32(gy)) == gy, "conversion must be exact"
    return gx, gy


def fmt(v):
    return "%.9g" % float(np.float32(v))


def dyadic_family():
    """-> list of (state cells dict {cell: [(x, y), ...]}, probe dict)"""
    out = []
    e = 2.0 ** -10
    # target offsets (dx, dy) from the proposal, per kind of neighbour
    axis = [1 - e, 1.0, 1 + e, 0.5, 1.5, 2.0]                    # along the direction's axis
    diag = [(0.75, 0.625), (0.75, 0.75), (0.625, 0.78125), (0.5, 0.5), (1.0, 1.0), (0.70703125, 0.70703125)]
    for (cx, cy) in [(1, 1), (0, 0), (3, 3), (0, 3), (2, 0)]:
        cell = cx + CPS * cy
        for hx in (-1, 0, 1):
            for hy in (-1, 0, 1):
                # proposal near the face / corner the direction points at; mover elsewhere in the cell
                px = 0.25 if hx < 0 else (2.25 if hx > 0 else 1.25)
                py = 0.25 if hy < 0 else (2.25 if hy > 0 else 1.25)
                offs = []
                if hx == 0 and hy == 0:
                    offs = [(r, 0.0) for r in axis[:5]] + [(0.0, -r) for r in axis[:3]] + diag[:3]
                elif hx != 0 and hy != 0:
                    offs = [(hx * a, hy * b) for a, b in diag]
                else:
                    offs = [(hx * r, hy * r) for r in axis]
                ncell = ((cx + hx) % CPS) + CPS * ((cy + hy) % CPS)
                for (ox, oy) in offs:
                    tx, ty = px + ox - hx * W, py + oy - hy * W         # target, local in its own cell
                    if not (0 < tx <= W and 0 < ty <= W):
                        continue
                    state = {}
                    own = [(1.25, 1.5 if (hx, hy) != (0, 0) else 2.25)]   # the mover's current position, slot 0
                    slot = 0
                    if hx == 0 and hy == 0:
                        # target shares the cell: mover in slot 1 so that j != i is exercised on both sides
                        own = [(tx, ty), own[0]]
                        slot = 1
                    else:
                        state[ncell] = [(tx, ty)]
                    # a distractor in another neighbour cell, in the attractive range of the proposal
                    dcell = ((cx + 1) % CPS) + CPS * cy if hx <= 0 else ((cx - 1) % CPS) + CPS * cy
                    state.setdefault(dcell, []).append((0.5, 1.25) if hx <= 0 else (2.0, 1.25))
                    state[cell] = own
                    out.append((state, dict(family="dyadic", cx=cx, cy=cy, slot=slot, px=px, py=py,
                                            own=own, note=f"dir=({hx},{hy}) off=({ox},{oy})")))
        # faces: closed [lb, ub] in the reference (subsweep.h:78-86), half-open (lb, ub] here
        for (px, py, note) in [(W, 1.0, "x upper face"), (0.0, 1.0, "x lower face"), (-e, 1.0, "x below"),
                               (W + e, 1.0, "x above"), (1.0, W, "y upper face"), (1.0, 0.0, "y lower face"),
                               (1.0, -e, "y below"), (1.0, W + e, "y above"), (W, W, "upper corner"), (e, e, "inside corner")]:
            own = [(1.0, 1.0)]
            out.append(({cell: own}, dict(family="dyadic", cx=cx, cy=cy, slot=0, px=px, py=py, own=own, note=note)))
    return out


def trajectory_family(n_M, delta, seed, sweeps, tag):
    o, phi = ref_geometry_oracle(64, n_M=n_M, move_delta=delta, seed=seed)
    disk, n = o.assign(o.init_r())
    episodes, probes = [], []
    for sweep in range(sweeps):
        order, f, d = o.schedule(sweep)
        for colour in order:
            off = o.colour_to_off(colour)
            ep = dict(tag=tag, n_M=n_M, move_delta=delta, seed=seed, phi=phi, sweep=sweep, off=off,
                      n=[int(v) for v in n],
                      disk=[[[fmt(disk[c, dim, s]) for s in range(n[c])] for dim in (0, 1)] for c in range(16)])
            acc0 = o.accepted.value
            o.trace_on(4 * n_M)
            o.subsweep(disk, n, off, sweep)
            rec = o.trace_off()
            ep["accepted"] = int(o.accepted.value - acc0)
            ep["n_after"] = [int(v) for v in n]
            ep["disk_after"] = [[[fmt(disk[c, dim, s]) for s in range(n[c])] for dim in (0, 1)] for c in range(16)]
            for r in rec:
                own = [(float(r.own_x[k]), float(r.own_y[k])) for k in range(r.cnt)]
                probes.append(dict(family="trajectory", episode=len(episodes), cx=r.cx, cy=r.cy, slot=r.slot,
                                   px=float(r.px), py=float(r.py), own=own, trial=r.trial, verdict=r.verdict))
            episodes.append(ep)
        o.shift_cells(disk, n, f, d)
    assert o.lost == 0
    return episodes, probes


def main():
    states, probes, episodes = [], [], []
    for state, p in dyadic_family():
        p["state"] = len(states)
        states.append(state)
        probes.append(p)
    for (n_M, delta, seed, sweeps, tag) in [(4, 0.2, 11, 4, "nM4"), (10, 0.1, 12, 2, "nM10")]:
        eps, prs = trajectory_family(n_M, delta, seed, sweeps, tag)
        base_ep, base_state = len(episodes), len(states)
        for ep in eps:
            st = {}
            for c in range(16):
                if ep["n"][c]:
                    st[c] = [(float(np.float32(ep["disk"][c][0][s])), float(np.float32(ep["disk"][c][1][s])))
                             for s in range(ep["n"][c])]
            states.append(st)
        for p in prs:
            p["state"] = base_state + p["episode"]
            p["episode"] += base_ep
        episodes += eps
        probes += prs
    # ---- harness input (reference conventions: global coordinates)
    lines = [str(len(states))]
    for st in states:
        lines.append(str(len(st)))
        for c in sorted(st):
            g = [to_global(c, x, y) for (x, y) in st[c]]
            lines.append(" ".join([str(c), str(len(g))] + [fmt(v[0]) for v in g] + [fmt(v[1]) for v in g]))
    lines.append(str(len(probes)))
    for p in probes:
        cell = p["cx"] + CPS * p["cy"]
        cxo, cyo = p["cx"] * W - LBOX / 2, p["cy"] * W - LBOX / 2
        gpx, gpy = cxo + p["px"], cyo + p["py"]
        assert float(np.float32(gpx)) == gpx and float(np.float32(gpy)) == gpy
        g = [to_global(cell, x, y) for (x, y) in p["own"]]
        lines.append(" ".join([str(p["state"]), str(p["cx"]), str(p["cy"]), str(p["slot"]), fmt(gpx), fmt(gpy),
                               str(len(g))] + [fmt(v[0]) for v in g] + [fmt(v[1]) for v in g]))
    open(os.path.join(HERE, "trial_probes_in.txt"), "w").write("\n".join(lines) + "\n")
    for p in probes:
        p["own"] = [[fmt(x), fmt(y)] for (x, y) in p["own"]]
        p["px"], p["py"] = fmt(p["px"]), fmt(p["py"])
    json.dump(dict(generator="tests/golden/make_trial_probes.py", geometry=dict(L=10, w=2.5, cps=4),
                   n_states=len(states), episodes=episodes, probes=probes),
              open(os.path.join(HERE, "trial_probes.json"), "w"), separators=(",", ":"))
    fam = {}
    for p in probes:
        fam[p["family"]] = fam.get(p["family"], 0) + 1
    print(f"{len(states)} states, {len(probes)} probes {fam}, {len(episodes)} episodes")


if __name__ == "__main__":
    main()
