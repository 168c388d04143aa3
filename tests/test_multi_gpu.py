"""Multi-GPU slab decomposition.

CPU part (gloo, world_size 2): the host-side logic every rank must agree on without
communication -- geometry, slab partition, per-sweep schedule -- plus the unique-id broadcast
plumbing.  GPU part: an R-rank NCCL run must be bit-identical to the single-GPU run (needs
>= 2 GPUs, skipped otherwise), and slab geometry invariants on one GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # 1. every rank derives the same geometry and schedule from (params, sweep): no broadcast needed
    o = O.Oracle(2 ** 16, phi=0.70, move_delta=0.1, cps_multiple=2 * world)
    sched = []
    for s in range(64):
        order, f, d = o.schedule(s)
        sched += order + [f, int(np.float32(d).view(np.uint32))]
    mine = torch.tensor([o.cps] + sched, dtype=torch.int64)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    # 2. slab partition: whole even row counts, contiguous, covering the box exactly once
    rows = o.cps // world
    row0 = rank * rows
    cover = torch.zeros(o.cps, dtype=torch.int64)
    cover[row0:row0 + rows] = 1
    dist.all_reduce(cover)
    part_ok = bool((cover == 1).all()) and rows % 2 == 0 and o.cps % (2 * world) == 0
    # 3. the 128-byte unique id travels from rank 0 to everybody (comm_init_from_torch plumbing)
    idt = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
    dist.broadcast(idt, src=0)
    id_ok = bool((idt == torch.arange(128, dtype=torch.uint8)).all())
    # 4. ring neighbours are mutual
    lower, upper = (rank + world - 1) % world, (rank + 1) % world
    nb = torch.tensor([lower, upper])
    allnb = [torch.empty_like(nb) for _ in range(world)]
    dist.all_gather(allnb, nb)
    ring_ok = all(int(allnb[int(allnb[r][1])][0]) == r for r in range(world))
    q.put((rank, same, part_ok, id_ok, ring_ok))
    dist.destroy_process_group()


def test_ranks_agree_without_communication_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert len(res) == 2
    for rank, same, part_ok, id_ok, ring_ok in res:
        assert same and part_ok and id_ok and ring_ok, (rank, same, part_ok, id_ok, ring_ok)


@pytest.mark.gpu
def test_slab_geometry_on_one_gpu(built):
    import pmc_b200
    world = 4
    total = 0
    for rank in range(world):
        mc = pmc_b200.ParallelMC(2 ** 16, phi=0.70, cps_multiple=2 * world, rank=rank, n_ranks=world)
        g = mc.geom
        assert g.rows * world == g.cps and g.row0 == rank * g.rows and g.ghost_rows == 5
        assert g.local_cells == (g.rows + 10) * g.cps
        # assign fills owned rows and ghost rows straight from r: no exchange needed
        disk, n = mc.assign(mc.init_r())
        own = n.view(-1, g.cps)[g.ghost_rows:g.ghost_rows + g.rows]
        total += int(own.sum())
        assert int(n.view(-1, g.cps)[:g.ghost_rows].sum()) > 0      # ghosts populated
        chk = mc.check(disk, n)
        assert chk["out_of_cell"] == 0 and chk["overlaps"] == 0
        mc.close()
    assert total == 2 ** 16


@pytest.mark.gpu
@pytest.mark.parametrize("N,sweeps,port,mode", [
    (2 ** 18, 7, 29611, "dense"),       # one launch per part and sweep
    (2 ** 22, 9, 29612, "dense"),       # 542 rows per rank: interior in 5 bands on 5 streams
    (2 ** 20, 12, 29613, "crowded"),    # cells with 7 / 8 disks in the ghost rows: the flags that travel with the ring
])
def test_two_rank_run_is_bit_identical_to_single_gpu(built, N, sweeps, port, mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2; log kept under profiles/r2/)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "scripts", "slab_worker.py"), str(N), str(sweeps)] + (["crowded"] if mode == "crowded" else [])
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "bit_identical=True" in out.stdout and "protocol_identical=True" in out.stdout
    if mode == "crowded":
        import re
        assert int(re.search(r"crowded_cells=(\d+)", out.stdout).group(1)) > 20
