"""Development aid: where does pmc_run_host spend its time (PCIe copies vs kernels)?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pmc_b200
N, S = 2 ** 24, int(os.environ.get("PMC_SWEEPS", 100))
mc = pmc_b200.ParallelMC(N, phi=0.70, move_delta=0.1, n_M=4)
g = mc.geom
r = mc.init_r()
r_host = r.cpu().pin_memory()
disk_host = torch.empty((g.local_cells, 2, 8), dtype=torch.float32).pin_memory()
n_host = torch.empty((g.local_cells,), dtype=torch.int16).pin_memory()
disk, n = mc.assign(r)
def timed(fn, reps=3):
    torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best * 1e3
print("H2D r      %.2f ms  (%.1f GB/s)" % ((t := timed(lambda: r.copy_(r_host, non_blocking=True))), r_host.numel() * 4 / t / 1e6))
print("D2H disk   %.2f ms  (%.1f GB/s)" % ((t := timed(lambda: disk_host.copy_(disk, non_blocking=True))), disk_host.numel() * 4 / t / 1e6))
print("D2H n      %.2f ms" % timed(lambda: n_host.copy_(n, non_blocking=True)))
print("assign     %.2f ms" % timed(lambda: mc.assign(r, disk, n)))
print("sweep x%d  %.2f ms" % (S, timed(lambda: mc.sweep(disk, n, 0, S))))
print("sweep x1   %.2f ms" % timed(lambda: mc.sweep(disk, n, 0, 1)))
print("run_host   %.2f ms" % timed(lambda: mc.run_host(r_host, 0, S, disk_host, n_host)))
