"""CPU, world_size = 2, gloo: the N>1 host logic (batch sharding, the single loss all-reduce, max-over-ranks
timing).  The per-shard losses come from the CPU oracle standing in for the kernels (no GPU here)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, B, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from learning3d_b200.dist import shard_batch, shard_bounds, global_mean_from_shards, max_over_ranks
    rng = np.random.default_rng(42)                       # identical full batch on every rank
    a = rng.random((B, 64, 3), dtype=np.float32)
    b = rng.random((B, 64, 3), dtype=np.float32)
    lo, hi = shard_bounds(B, rank, world)
    a_loc = shard_batch(torch.from_numpy(a)).numpy()
    assert a_loc.shape[0] == hi - lo and np.array_equal(a_loc, a[lo:hi])
    local = torch.tensor(oracle.chamfer_loss(a[lo:hi], b[lo:hi]), dtype=torch.float32)
    glob = global_mean_from_shards(local, hi - lo)
    slow = max_over_ranks(1.0 + rank, "cpu")
    # kNN shards are independent: concatenating per-rank results equals the full-batch result
    x = rng.random((B, 3, 48), dtype=np.float32)
    idx_loc = torch.from_numpy(oracle.knn_expansion(x[lo:hi], 5))
    gathered = [torch.empty((shard_bounds(B, r, world)[1] - shard_bounds(B, r, world)[0], 48, 5), dtype=torch.int64)
                for r in range(world)]
    dist.all_gather(gathered, idx_loc) if B % world == 0 else None
    if rank == 0:
        full = oracle.chamfer_loss(a, b)
        ok_knn = True
        if B % world == 0:
            ok_knn = np.array_equal(torch.cat(gathered).numpy(), oracle.knn_expansion(x, 5))
        out_q.put((float(glob), float(full), slow, ok_knn))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 5])     # 5: ragged shards (3 + 2)
def test_two_rank_sharding_and_loss_allreduce(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    glob, full, slow, ok_knn = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert abs(glob - full) < 1e-6          # shard means re-weighted by shard size == global mean
    assert slow == 2.0                      # max over ranks
    assert ok_knn


def test_shard_bounds_cover_everything():
    from learning3d_b200.dist import shard_bounds
    for total in (0, 1, 7, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
