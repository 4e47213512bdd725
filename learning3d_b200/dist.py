"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed.

Every kernel of the path is independent per batch item, so the path shards over the batch with NO
data-path collective: each rank runs the kernels on its contiguous slice.  The only exchange is the
scalar loss (and, in training, the parameter gradients, which belong to the caller's DDP wrapper):
Chamfer / EMD means are over B*N points, so shard means are re-weighted by shard size.
The reference has no distributed code at all (nn.DataParallel in examples/train_flownet.py:243-245).
"""
import torch
import torch.distributed as dist


def shard_bounds(total, rank, world):
    """Contiguous, balanced [lo, hi) slice of `total` batch items owned by `rank`."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world: %r/%r" % (rank, world))
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(t, rank=None, world=None):
    """The rank's slice of a batch-major tensor (dim 0)."""
    if rank is None:
        rank = dist.get_rank()
    if world is None:
        world = dist.get_world_size()
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def global_mean_from_shards(local_mean, local_items, group=None):
    """Combine per-rank means of a per-item quantity into the global mean with ONE all-reduce.

    local_mean: 0-d tensor, the mean over this rank's `local_items` batch items (e.g. the fused
    Chamfer loss of the shard); shards may be ragged.  Returns sum_r(mean_r * items_r) / sum_r(items_r).
    """
    buf = torch.stack([local_mean.detach().to(torch.float64) * float(local_items),
                       torch.tensor(float(local_items), dtype=torch.float64, device=local_mean.device)])
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return (buf[0] / buf[1]).to(local_mean.dtype)


def max_over_ranks(value, device, group=None):
    """max over ranks of a python float (used for device-side timings: the slowest rank defines a step)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
