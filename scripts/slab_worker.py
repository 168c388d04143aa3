"""torchrun worker: N-rank slab run == single-GPU run, bit for bit (SURVEY.md 8e test).
usage: torchrun --nproc-per-node R scripts/slab_worker.py [N] [sweeps] [crowded]
`crowded`: small disks thrown at random (Poisson occupancy, mean 1.5): cells with 7 and 8 disks sit in the
ghost rows too, so the crowded-cell flags that travel with the ring decide which kernel path a boundary
tile takes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import pmc_b200


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 2 ** 18
    sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    crowded = len(sys.argv) > 3 and sys.argv[3] == "crowded"
    kw = dict(phi=0.70, move_delta=0.1, n_M=4, seed=1234, cps_multiple=2 * world)
    if crowded:
        import numpy as np
        sigma, lam = 0.25, 1.5
        kw.update(sigma_d=sigma, phi=float(lam * np.pi * sigma * sigma / 16.0), cell_w=2.0, move_delta=0.3)
    mc = pmc_b200.ParallelMC(N, device=local, rank=rank, n_ranks=world, **kw)
    mc.strict = not crowded
    mc.comm_init_from_torch()
    g = mc.geom

    def start():
        if not crowded:
            return mc.init_r()
        rng = np.random.default_rng(12)                  # the same points on every rank
        hl = np.float32(g.L / 2)
        return torch.from_numpy((rng.random((2, N), dtype=np.float32) * 2 - 1) * hl * np.float32(0.9999)).cuda()
    disk, n = mc.assign(start())
    mc.sweep(disk, n, 0, sweeps)
    # per-call protocol on a second copy (sub-sweep kernel + stand-alone shift in slab mode)
    d2, n2 = mc.assign(start())
    for s in range(sweeps):
        order, f, d = mc.schedule(s)
        for c in order:
            mc.subsweep(d2, n2, mc.colour_to_off(c), s)
        mc.shift_cells(d2, n2, f, d)
    torch.cuda.synchronize()
    same_protocol = bool(torch.equal(disk.view(torch.int32), d2.view(torch.int32)) and torch.equal(n, n2))
    c = mc.counters()
    cnt = torch.tensor([c["trials"], c["accepted"], c["lost"], c["status"]], dtype=torch.float64, device="cuda")
    dist.all_reduce(cnt)
    G, rows, cps = g.ghost_rows, g.rows, g.cps
    own_d = disk.view(-1, cps, 2, 8)[G:G + rows].contiguous()
    own_n = n.view(-1, cps)[G:G + rows].to(torch.int32).contiguous()   # NCCL has no int16
    gd = [torch.empty_like(own_d) for _ in range(world)] if rank == 0 else None
    gn = [torch.empty_like(own_n) for _ in range(world)] if rank == 0 else None
    dist.gather(own_d, gd, dst=0)
    dist.gather(own_n, gn, dst=0)
    ok = torch.tensor([1 if same_protocol else 0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        full_d, full_n = torch.cat(gd).view(-1, 2, 8), torch.cat(gn).view(-1).to(torch.int16)
        ref = pmc_b200.ParallelMC(N, device=local, **kw)
        ref.strict = not crowded
        rd, rn = ref.assign(start())
        ref.sweep(rd, rn, 0, sweeps)
        rc = ref.counters()
        # fused sweep and the per-call protocol both count the same trials: the slab run did them twice
        same = torch.equal(full_d.view(torch.int32), rd.view(torch.int32)) and torch.equal(full_n, rn)
        counters = (cnt[0].item() == 2 * rc["trials"] and cnt[1].item() == 2 * rc["accepted"] and (crowded or cnt[3].item() == 0))
        crowded_cells = int((full_n >= 7).sum())
        print(f"SLAB ranks={world} N={N} cps={cps} rows/rank={rows} sweeps={sweeps} crowded_cells={crowded_cells} "
              f"bit_identical={same} protocol_identical={bool(ok.item())} counters_ok={counters}", flush=True)
        if not (same and ok.item() and counters):
            dist.destroy_process_group()
            sys.exit(1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
