"""N>1 host logic on CPU: phase ranges tile the schedule exactly and the rank-ordered checksum fold is what a
single process computes. world_size 2 over gloo (127.0.0.1)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from poppy_b200 import shard


def test_phase_ranges_tile_exactly():
    for n in (1, 7, 600, 2400, 2401):
        for world in (1, 2, 3, 4, 8):
            r = [shard.phase_range(k, world, n) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_phase_schedule_endpoints():
    s = shard.phase_schedule(600)
    assert s[0] == 0.0 and s[-1] == 1.0 and s.dtype == np.float32 and len(s) == 600
    assert (np.diff(s) > 0).all()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    lo, hi = shard.phase_range(rank, world, n_total)
    sched = shard.phase_schedule(n_total)[lo:hi]
    # stand-in for the per-rank frame checksum: a deterministic function of the rank's phases
    local = int(np.frombuffer(sched.tobytes(), np.uint8).astype(np.uint64).sum() * 2654435761 % (1 << 62))
    gathered = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, torch.tensor([local], dtype=torch.int64))
    t = torch.tensor([float(hi - lo)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    tmax = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        q.put((shard.combine_checksums([int(g.item()) for g in gathered]), t.item(), tmax.item()))
    dist.destroy_process_group()


def test_two_rank_partition_and_checksum_fold():
    n_total, world = 601, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, total, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sched = shard.phase_schedule(n_total)
    sums = []
    for r in range(world):
        lo, hi = shard.phase_range(r, world, n_total)
        sums.append(int(np.frombuffer(sched[lo:hi].tobytes(), np.uint8).astype(np.uint64).sum() * 2654435761 % (1 << 62)))
    assert got == shard.combine_checksums(sums)
    assert total == n_total and tmax == world
