"""Ad-hoc timing of the fused sweep (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import pmc_b200

def run(N, phi, sweeps, delta=0.1, n_M=4):
    mc = pmc_b200.ParallelMC(N, phi=phi, move_delta=delta, n_M=n_M)
    disk, n = mc.assign(mc.init_r())
    mc.set_blocking(0)
    mc.sweep(disk, n, 0, 10)
    torch.cuda.synchronize()
    mc.reset_counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mc.sweep(disk, n, 10, sweeps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    c = mc.counters()
    print(f"N={N} phi={phi} cps={mc.geom.cps} sweeps={sweeps} ms/sweep={ms/sweeps:.3f} "
          f"moves/s={c['trials']/ms*1e3:.3e} acc={c['accepted']/max(c['trials'],1):.3f} status={c['status']}", flush=True)

if __name__ == "__main__" and not os.environ.get("PMC_NOMAIN"):
    run(2**20, 0.70, 50)
    run(2**24, 0.70, 20)
    run(2**22, 0.30, 20, delta=0.4)
