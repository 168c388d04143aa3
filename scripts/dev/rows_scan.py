"""Development aid: us / sweep at N = 2^20 against tile_rows x bands (final kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, pmc_b200
mc = pmc_b200.ParallelMC(2 ** 20, phi=0.70, move_delta=0.1, n_M=4)
disk, n = mc.assign(mc.init_r())
mc.set_blocking(0)
mc.sweep(disk, n, 0, 300)
s = 300
for rows in (0, 22, 20, 18, 16, 14, 12):
    for bands in (4, 5, 6):
        mc.set_tuning("tile_rows", rows); mc.set_tuning("bands", bands)
        mc.sweep(disk, n, s, 300); s += 300
        torch.cuda.synchronize(); mc.reset_counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); mc.sweep(disk, n, s, 3000); e1.record(); torch.cuda.synchronize(); s += 3000
        ms = e0.elapsed_time(e1); c = mc.counters()
        print(f"tile_rows={rows} bands={bands} us/sweep={1e3*ms/3000:.2f} moves/s={c['trials']/ms*1e3:.3e}", flush=True)
