"""Multi-GPU plumbing of the hot path: one process per GPU, batch sharding, and the one collective.

Trajectories are independent, so the batch is cut into contiguous per-rank slices and every
rank runs the same kernels on its slice with NO data-path communication.  The only exchange is the
final "best trajectory" gather: each rank reduces its slice to (min cost, global index) and the pairs
are all-gathered (16 bytes per rank) -- `torch.distributed` over NCCL on GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(total, rank, world_size):
    """Contiguous slice [lo, hi) of `total` items for `rank`; sizes differ by at most one."""
    base, extra = divmod(int(total), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(array, rank=None, world_size=None, axis=0):
    """This rank's slice of a host array / tensor along `axis`."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    lo, hi = shard_bounds(array.shape[axis], rank, world_size)
    idx = [slice(None)] * array.ndim
    idx[axis] = slice(lo, hi)
    return array[tuple(idx)], lo


def gather_best(cost, offset=0, group=None):
    """Global (min cost, global index, owner rank) over all ranks' `cost [B_local]` tensors.

    Local part: one reduction on the device; exchange: one all_gather of a 2-element fp64 key.
    Ties go to the lowest global index (what np.argmin over the concatenated batch returns)."""
    cost = cost.detach()
    if cost.numel():
        cmin, imin = torch.min(cost.to(torch.float64), dim=0)
        key = torch.stack([cmin, (imin + offset).to(torch.float64)])
    else:
        key = torch.tensor([float("inf"), -1.0], dtype=torch.float64, device=cost.device)
    rank, w = world()
    if w == 1:
        return float(key[0]), int(key[1]), 0
    keys = [torch.empty_like(key) for _ in range(w)]
    dist.all_gather(keys, key, group=group)
    vals = [(float(k[0]), int(k[1]), r) for r, k in enumerate(keys) if int(k[1]) >= 0]
    return min(vals, key=lambda v: (v[0], v[1]))


def gather_costs(cost, group=None):
    """All ranks' per-trajectory costs concatenated in global order (optional, KBs)."""
    rank, w = world()
    if w == 1:
        return cost
    sizes = [torch.zeros(1, dtype=torch.int64, device=cost.device) for _ in range(w)]
    dist.all_gather(sizes, torch.tensor([cost.numel()], dtype=torch.int64, device=cost.device), group=group)
    n = max(int(s) for s in sizes)
    pad = torch.zeros(n, dtype=cost.dtype, device=cost.device)
    pad[: cost.numel()] = cost
    out = [torch.empty_like(pad) for _ in range(w)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: int(s)] for o, s in zip(out, sizes)])
