#!/usr/bin/env python
"""bench.py -- hard-disk trial moves/sec (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W            # ours, N=2^24 phi=0.70 on one B200
    torchrun ... bench.py --gpus N --steps K --warmup W      # ours, N=2^28 slabs over N B200s
    python bench.py --impl reference ...                     # the CPU path (oracle, all host cores)

A "step" is one pass of the hot path over one batch: `sweeps_per_step` full MC sweeps
(each = 4 checkerboard sub-sweeps + grid shift, start.cu:237-260) of the resident system.
value  = trial moves actually executed (device-counted) / device time, inputs resident in HBM.
e2e    = the same through pmc_run_host with HOST buffers: H2D(r) + assign + sweeps + D2H(disk, n) per step,
         steps being independent jobs alternated between two handles / two slab copies so that the copies of one
         job overlap the sweeps of the other.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N, phi, cps_multiple, move_delta)
    "n16m_phi0.70": (2 ** 24, 0.70, 2, 0.1),       # BASELINE metric config (1 GPU)
    "n256m_phi0.70": (2 ** 28, 0.70, 16, 0.1),     # BASELINE config 4 (1/2/4/8 GPUs)
    "n1m_phi0.70": (2 ** 20, 0.70, 2, 0.1),        # config 2
    "n16m_phi0.716": (2 ** 24, 0.716, 2, 0.1),     # config 3
    "n4m_phi0.30": (2 ** 22, 0.30, 2, 0.4),        # config 5 (dilute)
    "n4096_phi0.70": (4096, 0.70, 2, 0.1),         # config 1
}
N_M, NMAX, SIGMA, CELL_W, SEED = 4, 8, 1.0, 2.0, 1234
METRIC = "hard-disk trial moves/sec"
UNIT = "moves/s"

PROFILE_ROUND = "r2"


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of the fused sweep kernel PER SWEEP, from the committed
    `ncu --set full` capture of this workload (profiles/<round>/traffic_<workload>.json names the command,
    the build and the launches it summed) - or None: nothing is typed in here."""
    path = os.path.join(ROOT, "profiles", PROFILE_ROUND, f"traffic_{workload}.json")
    if not os.path.exists(path):
        return None, None
    t = json.load(open(path))
    return float(t["dram_bytes_per_sweep"]), os.path.relpath(path, ROOT)


def default_workload(world_env):
    """BASELINE.json metric: N=16M phi=0.70 on 1 GPU; N=256M at 1/2/4/8 GPUs.  The scaling series is the one
    launched through torch.distributed.run (WORLD_SIZE set, also for N=1), so all four lines share the 256M
    workload and the driver's efficiency is 256M against 256M."""
    return "n256m_phi0.70" if world_env else "n16m_phi0.70"


def bind_to_gpu_numa(torch, device):
    """Slab runs copy 0.6 GB per rank and step each way: pin this process (and therefore the first-touch placement
    of its pinned buffers) to the CPUs of the NUMA node the GPU hangs off.  Returns what was done (for the JSON)."""
    try:
        pr = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        cpulist = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "no local cpus in the affinity mask"
        os.sched_setaffinity(0, cpus)
        return f"{bdf}: {len(cpus)} local cpus"
    except Exception as e:                      # sysfs not exposed (containers): leave the mask alone
        return f"unchanged ({type(e).__name__})"


def state_hash(torch, disk, n, g, dist):
    """64-bit hash of the owned cells (position bits and counts), a wrapping SUM over cells of a mix of
    (global cell id, the cell's 16 words, its count): independent of how cells are spread over ranks, so
    the all-reduced value of a k-GPU run equals the 1-GPU value iff the states are bit-identical."""
    M1, M2 = -7046029254386353131, -4658895280553007687         # odd 64-bit constants (as int64)
    owned0 = g.ghost_rows * g.cps
    ncell = g.rows * g.cps
    w = torch.tensor([(2 * k + 1) * 0x9E3779B1 for k in range(1, 17)], dtype=torch.int64, device=disk.device)
    total = torch.zeros((), dtype=torch.int64, device=disk.device)
    step = 1 << 22
    for c0 in range(0, ncell, step):
        c1 = min(ncell, c0 + step)
        words = disk[owned0 + c0:owned0 + c1].reshape(c1 - c0, 16).view(torch.int32).to(torch.int64)
        h = (words * w).sum(dim=1) + n[owned0 + c0:owned0 + c1].to(torch.int64) * 0x632BE5AB
        gid = torch.arange(c0, c1, dtype=torch.int64, device=disk.device) + g.row0 * g.cps
        h = (h ^ (gid * M1)) * M2
        h = h ^ (h >> 29)
        total += (h * M1).sum()
    if dist:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    return "%016x" % (int(total.item()) & 0xFFFFFFFFFFFFFFFF)


def algorithmic_bytes_per_sweep(n_particles, n_cells):
    """56 B per particle per sweep + 12 B per cell per sweep."""
    return 56.0 * n_particles + 12.0 * n_cells


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(workload, sweeps, omp=True):
    """The reference has no CPU path; the CPU baseline is our C restatement
    (oracle/, kind="port"), OpenMP over same-colour cells on all host cores."""
    from oracle import oracle as O
    N, phi, mult, delta = WORKLOADS[workload]
    o = O.Oracle(N, phi=phi, sigma_d=SIGMA, cell_w=CELL_W, nmax=NMAX, n_M=N_M, move_delta=delta,
                 seed=SEED, cps_multiple=mult)
    disk, n = o.assign(o.init_r())
    O.set_threads()                                 # every core of the affinity mask, whatever OMP_NUM_THREADS says
    t0 = time.perf_counter()
    threads = o.sweep(disk, n, 0, sweeps, omp=omp)
    dt = time.perf_counter() - t0
    return o.trials.value / dt, threads, dt, o


def run_reference(args, rank, world):
    """--impl reference: the CPU path on the box's host cores, rank 0 only."""
    if rank != 0:
        return
    workload = args.workload or default_workload("WORLD_SIZE" in os.environ or args.gpus > 1)
    sample_wl = workload if WORKLOADS[workload][0] <= 2 ** 24 else "n16m_phi0.70"
    sweeps = args.ref_sweeps
    from oracle import oracle as O
    O.set_threads()             # torchrun exports OMP_NUM_THREADS=1: use the affinity mask instead
    N, phi, mult, delta = WORKLOADS[sample_wl]
    o = O.Oracle(N, phi=phi, sigma_d=SIGMA, cell_w=CELL_W, nmax=NMAX, n_M=N_M, move_delta=delta,
                 seed=SEED, cps_multiple=mult)
    disk, n = o.assign(o.init_r())
    threads = 1
    for w in range(min(args.warmup, 1)):
        threads = o.sweep(disk, n, 0, 1, omp=True)
    tr0 = o.trials.value
    t0 = time.perf_counter()
    for k in range(args.steps):
        threads = o.sweep(disk, n, 1 + k * sweeps, sweeps, omp=True)
    dt = time.perf_counter() - t0
    v = (o.trials.value - tr0) / dt
    sample = f"{sample_wl}: {sweeps} sweeps/step from the lattice start, OpenMP over same-colour cells on {threads} threads"
    if sample_wl != workload:
        sample += (f"; the {workload} workload itself is NOT run on the CPU: the sample is the 16M system "
                   "(moves/s of this path does not depend on N once the state exceeds the caches)")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong" if workload.startswith("n256m") else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "sample_workload": sample_wl, "n_M": N_M, "nmax": NMAX,
                   "cell_w": o.g.w, "move_delta": delta, "sweeps_per_step": sweeps,
                   "proposal": "uniform square"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_JSON_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout: everything else that a library writes to file descriptor 1
    (NCCL prints its version banner there when NCCL_DEBUG is set) is sent to stderr instead."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--sweeps-per-step", type=int, default=1000,
                    help="MC sweeps per step; default = the reference's MCpasses (start.cu:24)")
    ap.add_argument("--burn-in", type=int, default=300)
    ap.add_argument("--ref-sweeps", type=int, default=20, help="sweeps per step of the CPU arm")
    ap.add_argument("--cpu-sweeps", type=int, default=60, help="sweeps of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import pmc_b200
    import __graft_entry__ as ge
    if rank == 0 and pmc_b200._stale():
        ge.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"
    n_ranks = world
    numa = bind_to_gpu_numa(torch, local_rank) if n_ranks > 1 else None

    workload = args.workload or default_workload("WORLD_SIZE" in os.environ or n_ranks > 1)
    N, phi, mult, delta = WORKLOADS[workload]
    if n_ranks > 1:
        mult = max(mult, 2 * n_ranks)
    W = max(args.warmup, 3)
    S = args.sweeps_per_step

    mc = pmc_b200.ParallelMC(N, phi=phi, sigma_d=SIGMA, cell_w=CELL_W, nmax=NMAX, n_M=N_M,
                             move_delta=delta, seed=SEED, cps_multiple=mult, device=local_rank,
                             rank=rank, n_ranks=n_ranks)
    if n_ranks > 1:
        mc.comm_init_from_torch()
    g = mc.geom
    # dense configurations: the reference's lattice (init_r) + burn-in; RSA jams at phi ~ 0.547
    # (SURVEY H6), so random sequential addition serves the dilute configuration only
    init = "rsa" if phi < 0.5 else "lattice"
    r = mc.rsa(seed=SEED) if init == "rsa" else mc.init_r()
    disk, n = mc.assign(r)
    del r
    torch.cuda.empty_cache()
    mc.set_blocking(0)
    sweep = 0
    mc.sweep(disk, n, sweep, args.burn_in)          # melt the lattice
    sweep += args.burn_in
    for _ in range(W):
        mc.sweep(disk, n, sweep, S)
        sweep += S
    torch.cuda.synchronize()
    mc.reset_counters()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        mc.sweep(disk, n, sweep, S)
        sweep += S
    e1.record()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    c = mc.counters()
    kernel_ms, kernel_launches = mc.kernel_time()   # CUDA events around the fused sweep kernels alone (per sweep)
    gpu_launches = mc.launch_count()                # kernels actually launched in the timed region
    trials = torch.tensor([c["trials"], c["accepted"], c["lost"], c["status"]], dtype=torch.float64, device="cuda")
    tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(trials, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    tot_trials, tot_acc, tot_lost, status = (float(x) for x in trials.tolist())
    value = tot_trials / (ms * 1e-3)
    fast = kernel_launches > 0
    n_sweeps_timed = args.steps * S

    # invariants after the timed region (cheap, device side)
    chk = mc.check(disk, n)
    chk_t = torch.tensor([chk["total"], chk["out_of_cell"], chk["overlaps"]], dtype=torch.float64, device="cuda")
    min_t = torch.tensor([chk["min_d2"]], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(chk_t, op=dist.ReduceOp.SUM)
        dist.all_reduce(min_t, op=dist.ReduceOp.MIN)
    total_particles = int(chk_t[0].item())
    min_d2 = float(min_t.item())
    # bit-identity across GPU counts: the same arguments give the same hash at every N
    hash_hex = state_hash(torch, disk, n, g, dist)
    end_sweep = sweep

    # roofline of the dominant kernel (the fused sweep kernel, one launch per sweep)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    if fast:
        # CUDA events bracket the sweep kernels of each pmc_sweep call (import / export excluded)
        ms_per_sweep = kernel_ms / kernel_launches
        kname = ("sweep4_kernel<4 CTAs/SM, fast> (one MC sweep = 4 colours + shiftCells; tile 24..30 x 24..28 cells "
                 "chosen per sweep; one sweep = several band launches of this kernel on their own streams)")
    else:
        ms_per_sweep = ms / n_sweeps_timed              # upper bound: includes 1 stand-alone shift per step
        kname = "sweep_tile_kernel<4,26,32,320,2,*> (generic path)"
    alg_bytes = algorithmic_bytes_per_sweep(N, g.n_cells) / n_ranks      # per sweep per GPU
    achieved = alg_bytes / (ms_per_sweep * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(workload) if n_ranks == 1 else (None, None)
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "traffic": traffic, "traffic_source": traffic_src,
                "launch_unit": "one MC sweep (all band launches of the fused kernel)",
                "algorithmic_bytes_per_launch": alg_bytes,
                "bytes_per_move": alg_bytes * n_ranks / (tot_trials / n_sweeps_timed),
                "ms_per_sweep": ms_per_sweep,
                "limiter": "instruction issue, not DRAM (ncu: profiles/%s); the HBM figure is the contract's roofline" % PROFILE_ROUND,
                "issue_ceiling_lane_instr_per_s": 148 * 128 * 1.965e9}

    # end to end through the C-ABI with host buffers (single GPU only: slabs keep state on device)
    e2e = None
    if not args.no_e2e and n_ranks == 1:
        # a stream of independent jobs through pmc_run_host (HOST buffers in, HOST buffers out), alternating between
        # two handles on two streams, so that one job's H2D / D2H overlap the other job's sweeps
        del disk, n
        torch.cuda.empty_cache()
        r_host = mc.init_r().cpu().pin_memory()
        outs = [(torch.empty((g.local_cells, 2, NMAX), dtype=torch.float32).pin_memory(),
                 torch.empty((g.local_cells,), dtype=torch.int16).pin_memory()) for _ in range(2)]
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        with torch.cuda.stream(streams[0]):
            mc.use_torch_stream()
        with torch.cuda.stream(streams[1]):
            mc_b = pmc_b200.ParallelMC(N, phi=phi, sigma_d=SIGMA, cell_w=CELL_W, nmax=NMAX, n_M=N_M,
                                       move_delta=delta, seed=SEED, cps_multiple=mult, device=local_rank)
        hs = [mc, mc_b]
        for m in hs:
            m.set_blocking(0)

        def run_jobs(k):
            for i in range(k):
                hs[i & 1].run_host(r_host, 0, S, *outs[i & 1])
            for m in hs:
                m.synchronize()

        run_jobs(2)
        for m in hs:
            m.reset_counters()
        torch.cuda.synchronize()
        f0 = torch.cuda.Event(enable_timing=True)
        f1 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        t0 = time.perf_counter()
        f0.record(streams[0])
        streams[1].wait_event(f0)
        run_jobs(args.steps)
        for b in range(2):
            f1[b].record(streams[b])
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        e2e_trials = sum(m.counters()["trials"] for m in hs)
        e2e_status = sum(m.counters()["status"] for m in hs)
        e2e_ms = max(f0.elapsed_time(f1[0]), f0.elapsed_time(f1[1]), wall * 1e3)
        same = bool(torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])) if args.steps > 1 else True
        last = outs[(args.steps - 1) & 1]
        e2e = {"value": e2e_trials / (e2e_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(r_host.numel() * 4),
               "d2h_bytes_per_step": int(last[0].numel() * 4 + last[1].numel() * 2),
               "ms_per_step": e2e_ms / args.steps,
               "what": "a stream of independent jobs; per step pmc_run_host: H2D(r) + assign + sweeps + D2H(disk, n), pinned "
                       "host buffers; jobs alternate between two handles on two streams (non-blocking pmc_run_host), so "
                       "the copies of one job overlap the sweeps of the other",
               "both_handles_same_result": same, "status": int(e2e_status),
               "n_sum_check": int(last[1].to(torch.int64).sum().item())}
    elif not args.no_e2e:
        # slab runs: a stream of independent jobs, like the single-GPU leg (every step starts from the same host
        # input).  Per step and rank: H2D(slab disk, n) from pinned memory + pmc_sweep (S sweeps, NCCL ghost rows)
        # + D2H(slab disk, n) into pinned memory.  Two device copies of the slab: the H2D of step i + 1 and the
        # D2H of step i - 1 run on copy streams while step i sweeps.  (Inside ONE step nothing can overlap: a band
        # of rows can run sweep t only when its neighbours have done sweep t - 1, so while the copy is in flight at
        # most ~bands/2 of the 1000 sweeps could start.)
        disk_in = disk.cpu().pin_memory()
        n_in = n.cpu().pin_memory()
        dbuf = [(disk, n), (torch.empty_like(disk), torch.empty_like(n))]
        hout = [(torch.empty_like(disk_in).pin_memory(), torch.empty_like(n_in).pin_memory()) for _ in range(2)]
        main_s = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        ev_h2d = [torch.cuda.Event() for _ in range(2)]
        ev_swept = [torch.cuda.Event() for _ in range(2)]
        ev_d2h = [torch.cuda.Event() for _ in range(2)]
        mc.set_blocking(0)

        def h2d(b):
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_d2h[b])          # the previous result of this buffer has left the device
                dbuf[b][0].copy_(disk_in, non_blocking=True); dbuf[b][1].copy_(n_in, non_blocking=True)
                ev_h2d[b].record(s_in)

        def run_steps(k, sweep0):
            for b in range(2):
                ev_d2h[b].record(main_s)
            h2d(0)
            for i in range(k):
                b = i & 1
                if i + 1 < k:
                    h2d(b ^ 1)
                main_s.wait_event(ev_h2d[b])
                mc.sweep(dbuf[b][0], dbuf[b][1], sweep0, S)
                ev_swept[b].record(main_s)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_swept[b])
                    hout[b][0].copy_(dbuf[b][0], non_blocking=True); hout[b][1].copy_(dbuf[b][1], non_blocking=True)
                    ev_d2h[b].record(s_out)
            main_s.wait_stream(s_out)

        run_steps(2, sweep)
        torch.cuda.synchronize()
        mc.reset_counters()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        run_steps(args.steps, sweep)
        f1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dist.barrier()
        ce = mc.counters()
        # every step ran the same job: its result must be the same bits (and the serial, unpipelined answer)
        same = bool(torch.equal(hout[0][0], hout[1][0]) and torch.equal(hout[0][1], hout[1][1])) if args.steps > 1 else True
        disk.copy_(disk_in); n.copy_(n_in)
        mc.set_blocking(1)
        mc.sweep(disk, n, sweep, S)
        same = same and bool(torch.equal(disk.cpu(), hout[(args.steps - 1) & 1][0]))
        et = torch.tensor([max(f0.elapsed_time(f1), wall * 1e3)], dtype=torch.float64, device="cuda")
        etr = torch.tensor([ce["trials"], 0.0 if same else 1.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
        dist.all_reduce(etr, op=dist.ReduceOp.SUM)
        e2e_ms = float(et.item())
        nbytes = int(disk_in.numel() * 4 + n_in.numel() * 2) * n_ranks
        e2e = {"value": float(etr[0].item()) / (e2e_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
               "ms_per_step": e2e_ms / args.steps,
               "results_identical_to_serial_run": bool(etr[1].item() == 0.0),
               "host_numa_binding": numa,
               "what": "a stream of independent jobs; per step and rank: H2D(slab disk, n) + pmc_sweep (NCCL ghost rows) + "
                       "D2H(slab disk, n), pinned host buffers, two device copies of the slab so that the copies of "
                       "neighbouring steps overlap the sweeps; max over ranks"}

    cpu = None
    if rank == 0 and n_ranks == 1 and not args.no_cpu_baseline:
        sample_wl = workload if N <= 2 ** 24 else "n16m_phi0.70"
        v, threads, dt, _ = cpu_baseline(sample_wl, args.cpu_sweeps)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{sample_wl}: {args.cpu_sweeps} sweeps from the lattice start in {dt:.1f} s; "
                         "oracle/pmc_oracle.c (the reference has no CPU path), OpenMP over same-colour cells"}

    failed = False
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_ranks, "steps": args.steps,
            "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if workload.startswith("n256m") else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload, "n_particles": N, "phi": phi, "cells_per_side": g.cps,
                       "cell_w": g.w, "nmax": NMAX, "n_M": N_M, "move_delta": delta,
                       "proposal": "uniform square", "sweeps_per_step": S, "burn_in_sweeps": args.burn_in,
                       "init": ("random sequential addition (pmc_rsa_host)" if init == "rsa" else "square lattice (init_r)") + " + burn-in", "seed": SEED,
                       "l2": "state (%.0f MB per GPU) larger than L2, no flush" % (g.local_cells * 66 / 1e6),
                       "parallelism": "1 GPU" if n_ranks == 1 else f"{n_ranks} slabs of {g.rows} cell rows, NCCL ghost-row ring"},
            "acceptance": tot_acc / tot_trials, "trials": tot_trials, "lost": tot_lost, "status": int(status),
            "invariants": {"particles": total_particles, "out_of_cell": int(chk_t[1].item()),
                           "overlaps_below_sigma": int(chk_t[2].item()), "min_d2": min_d2,
                           "state_hash": hash_hex, "state_hash_after_sweeps": int(end_sweep),
                           "what": "hash = wrapping 64-bit sum over owned cells of mix(global cell id, position bits, count), "
                                   "all-reduced: equal at every GPU count iff the states are bit-identical"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": int(gpu_launches),
        }
        emit(line)
        # a run that dropped a disk, produced an overlap or lost a particle is not a measurement
        bad = [k for k, v in (("status", int(status)), ("lost", tot_lost), ("out_of_cell", chk_t[1].item()),
                              ("overlaps_below_sigma", chk_t[2].item()), ("particles_missing", total_particles - N)) if v]
        if e2e and e2e.get("results_identical_to_serial_run") is False:
            bad.append("pipelined e2e result differs from the serial run")
        if e2e and (e2e.get("both_handles_same_result") is False or e2e.get("status") or e2e.get("n_sum_check", N) != N):
            bad.append("e2e jobs disagree / lost particles")
        if bad:
            print("bench.py: invariant violated: " + ", ".join(bad), file=sys.stderr)
            failed = True
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    if failed:
        raise SystemExit(3)


if __name__ == "__main__":
    main()
