"""Development aid: ms / sweep of the fused kernel for small systems against the tile_rows knob."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, pmc_b200
def run(N, phi, delta, rows, bands, sweeps=3000):
    mc = pmc_b200.ParallelMC(N, phi=phi, move_delta=delta, n_M=4)
    if rows: mc.set_tuning("tile_rows", rows)
    if bands: mc.set_tuning("bands", bands)
    r = mc.rsa(seed=1234) if phi < 0.5 else mc.init_r()
    disk, n = mc.assign(r)
    mc.set_blocking(0)
    mc.sweep(disk, n, 0, 300); torch.cuda.synchronize(); mc.reset_counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); mc.sweep(disk, n, 300, sweeps); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); c = mc.counters()
    print(f"N={N} phi={phi} tile_rows={rows} bands={bands} us/sweep={1e3*ms/sweeps:.2f} moves/s={c['trials']/ms*1e3:.3e} status={c['status']}", flush=True)
for rows in (0, 20, 16, 12, 8):
    for bands in (6, 8):
        run(2**20, 0.70, 0.1, rows, bands)
for rows in (0, 16, 12):
    run(2**22, 0.30, 0.4, rows, 6, 1500)
run(2**24, 0.70, 0.1, 0, 6, 300)
