"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

Only the parts of the path that shard are here:

* independent networks (BASELINE config 5, temperature sweeps): `shard_range` gives every rank a contiguous block of
  networks; no data-path collective, results are gathered at the end (`gather_rates`);
* phase 2 (BASELINE config 4): the per-a reductions are independent once the reduced A ∪ B block exists.  Every rank
  runs `ngt_phase_two_shard(rank, world)` on every world-th a and the (1-P'_aa, tau'_a) pairs are merged with ONE
  all-reduce(sum) (unowned entries are zero) and installed with `ngt_set_final`.  Optionally rank 0's reduced block is
  broadcast first so that every rank reduces bit-identical input.

The reference has nothing to compare with here: it is single-threaded ("Parallelism strategies: none").
"""
import numpy as np


def shard_range(n_items, rank, world):
    """contiguous block [lo, hi) of `n_items` for `rank` (sizes differ by at most one)"""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist
    return dist


def all_reduce_sum(array, group=None):
    """sum a float64 numpy array over the ranks of `group` (moves through the GPU when the backend is NCCL)"""
    import torch
    dist = _dist()
    t = torch.from_numpy(np.ascontiguousarray(array, np.float64).copy())
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


class _DeviceBlock(object):
    """exposes a raw device pointer through __cuda_array_interface__ so torch can broadcast it in place"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 3}


def broadcast_reduced_block(handle, src=0, group=None):
    """NCCL broadcast of the phase-1-reduced (|A|+|B|) x ld block (incl. tau column) from rank `src` into every rank's
    engine memory (8 MiB at |A|+|B| = 1024)."""
    import torch
    ptr, n = handle.reduced_block_dev()
    t = torch.as_tensor(_DeviceBlock(ptr, n), device="cuda")
    _dist().broadcast(t, src=src, group=group)
    torch.cuda.synchronize()


def phase_two_sharded(handle, rank, world, reduce_fn=all_reduce_sum):
    """per-a reductions of this rank's share + one all-reduce; afterwards every rank's handle returns the full rates.
    `handle` needs phase_two_shard(rank, world), final() -> (om, tau) and set_final(om, tau)."""
    handle.phase_two_shard(rank, world)
    om, tau = handle.final()
    m = om.size
    merged = reduce_fn(np.concatenate([om, tau]))
    handle.set_final(merged[:m], merged[m:])
    return merged[:m], merged[m:]


def compute_rates_sharded(handle, group=None, broadcast=False):
    """phase 1 on every rank (replicated), optional broadcast of rank 0's reduced block, sharded phase 2."""
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    handle.phase_one()
    if broadcast and world > 1:
        broadcast_reduced_block(handle, 0, group)
    return phase_two_sharded(handle, rank, world, lambda a: all_reduce_sum(a, group))


def gather_rates(local_rates, n_total, rank, world, group=None):
    """all-gather per-network rate rows computed on contiguous shards (uses an all-reduce of a zero-padded array)"""
    local_rates = np.asarray(local_rates, np.float64)
    lo, hi = shard_range(n_total, rank, world)
    full = np.zeros((n_total,) + local_rates.shape[1:])
    full[lo:hi] = local_rates
    return all_reduce_sum(full, group)


# ---------------------------------------------------------------------------------------------------------------
# One dense network over several GPUs (BASELINE config 3 at N > 1): 1-D block-cyclic columns + panel broadcast.
# The engine exposes the per-block stages (include/ngt_b200.h, ngt_dist_*); the collectives are issued here.
# ---------------------------------------------------------------------------------------------------------------
def _dev_tensor(ptr, n):
    import torch
    return torch.as_tensor(_DeviceBlock(ptr, n), device="cuda")


def dense_phase_one_distributed(handle, group=None):
    """Phase 1 of one complete network spread over the ranks of `group` (every rank holds a handle created from the
    same rate matrix).  Per outer block: all-reduce of the block rows' outside mass (for the subtraction-free
    pivots), panel factorisation on the owning rank, NCCL broadcast of the factored panel, trailing update of the
    own columns on every rank.  Afterwards handle.phase_two() / rates() work as usual on every rank.

    The engine runs the bulk of each trailing update on its own second stream while the next block's collectives
    and panel proceed, so only the stream the collectives were issued on is synchronised here — never the device."""
    import torch
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    cur = torch.cuda.current_stream()
    handle.dist_begin(rank, world)
    npiv, block = handle.dist_info()
    for J in range(0, npiv, block):
        p, n = handle.dist_rowsum(J)
        dist.all_reduce(_dev_tensor(p, n), op=dist.ReduceOp.SUM, group=group)
        cur.synchronize()             # the engine works on its own streams: the collective must have landed
        owner = (J // block) % world
        if rank == owner:
            handle.dist_factor(J)
        p, n = handle.dist_panel(J)
        dist.broadcast(_dev_tensor(p, n), src=dist.get_global_rank(group, owner) if group is not None else owner, group=group)
        cur.synchronize()
        handle.dist_apply(J)
    p, n = handle.dist_root()
    dist.all_reduce(_dev_tensor(p, n), op=dist.ReduceOp.SUM, group=group)
    cur.synchronize()
    handle.dist_finish()


def dense_phase_one_emulated(handles):
    """The same schedule with the ranks emulated by `handles` inside ONE process on ONE GPU (tests): the collectives
    become explicit sums / copies between the handles' device buffers."""
    import torch
    world = len(handles)
    for g, h in enumerate(handles):
        h.dist_begin(g, world)
    npiv, block = handles[0].dist_info()
    for J in range(0, npiv, block):
        parts = [_dev_tensor(*h.dist_rowsum(J)) for h in handles]
        total = torch.stack(parts).sum(0)
        for t in parts:
            t.copy_(total)
        torch.cuda.synchronize()
        owner = (J // block) % world
        handles[owner].dist_factor(J)
        src = _dev_tensor(*handles[owner].dist_panel(J))
        for g, h in enumerate(handles):
            if g != owner:
                _dev_tensor(*h.dist_panel(J)).copy_(src)
        torch.cuda.synchronize()
        for h in handles:
            h.dist_apply(J)
    parts = [_dev_tensor(*h.dist_root()) for h in handles]
    total = torch.stack(parts).sum(0)
    for t in parts:
        t.copy_(total)
    torch.cuda.synchronize()
    for h in handles:
        h.dist_finish()
