"""Multi-GPU plumbing (one process per GPU).

Only the parts of the path that shard are here:

* independent networks (BASELINE config 5, temperature sweeps): `shard_range` gives every rank a contiguous block of
  networks; no data-path collective, the per-network rates are gathered at the end (`gather_rates`);
* phase 2 (BASELINE config 4): every rank runs its share of the phase-2 elimination tree
  (`ngt_phase_two_shard(rank, world)`: a contiguous block of every group's members) and the (1-P'_aa, tau'_a)
  vectors are merged with ONE all-reduce(sum) (entries of other ranks are zero).  On GPUs the all-reduce and the
  optional broadcast of rank 0's reduced block are NCCL calls made by the C++ engine on its own stream
  (`compute_rates_sharded` -> `ngt_phase_two_sharded`); `phase_two_sharded` is the same schedule with the reduction
  supplied by the caller (gloo in the CPU tests);
* one dense network over several GPUs (BASELINE config 3 at N > 1): `dense_phase_one_distributed` ->
  `ngt_dist_phase_one`, the whole block loop in C++ with NCCL on the engine's streams.

torch.distributed is used for rendezvous only (handing NCCL's unique id to the ranks, gathering a few numbers).
The reference has nothing to compare with here: it is single-threaded ("Parallelism strategies: none").
"""
import numpy as np

from . import _capi


def shard_range(n_items, rank, world):
    """contiguous block [lo, hi) of `n_items` for `rank` (sizes differ by at most one, remainder to the low ranks;
    the same split ngt_phase_two_shard uses for the members of A and of B)"""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist
    return dist


_comms = {}


def communicator(group=None, device=None):
    """the library's NCCL communicator for `group` (created once per process and group)"""
    import torch
    if device is None:
        device = torch.cuda.current_device()
    key = (id(group) if group is not None else None, int(device))
    if key not in _comms:
        _comms[key] = _capi.Comm.from_torch(device=device, group=group)
    return _comms[key]


def all_reduce_sum(array, group=None):
    """sum a small float64 numpy array over the ranks of `group` (host-side plumbing: gathers of a few rates)"""
    import torch
    dist = _dist()
    t = torch.from_numpy(np.ascontiguousarray(array, np.float64).copy())
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def phase_two_sharded(handle, rank, world, reduce_fn=all_reduce_sum):
    """this rank's share of phase 2 + one all-reduce supplied by the caller; afterwards every rank's handle returns the
    full rates.  `handle` needs phase_two_shard(rank, world), final() -> (om, tau) and set_final(om, tau)."""
    handle.phase_two_shard(rank, world)
    om, tau = handle.final()
    shape = np.shape(om)
    merged = reduce_fn(np.concatenate([np.ravel(om), np.ravel(tau)]))
    half = merged.size // 2
    om, tau = merged[:half].reshape(shape), merged[half:].reshape(shape)
    handle.set_final(om, tau)
    return om, tau


def compute_rates_sharded(handle, group=None, broadcast=False, comm=None):
    """phase 1 on every rank (replicated), optional NCCL broadcast of rank 0's reduced block, sharded phase 2 with the
    merge done by NCCL on the engine's stream.  Returns the device time of the phase-2 part in ms."""
    comm = comm or communicator(group)
    handle.phase_one()
    return handle.phase_two_sharded(comm, broadcast_root=broadcast)


def gather_rates(local_rates, n_total, rank, world, group=None):
    """all-gather per-network rate rows computed on contiguous shards (uses an all-reduce of a zero-padded array)"""
    local_rates = np.asarray(local_rates, np.float64)
    lo, hi = shard_range(n_total, rank, world)
    full = np.zeros((n_total,) + local_rates.shape[1:])
    full[lo:hi] = local_rates
    return all_reduce_sum(full, group)


# ---------------------------------------------------------------------------------------------------------------
# One dense network over several GPUs (BASELINE config 3 at N > 1): 1-D block-cyclic columns + panel broadcast.
# ---------------------------------------------------------------------------------------------------------------
def dense_phase_one_distributed(handle, group=None, comm=None):
    """Phase 1 of one complete network spread over the ranks of `group` (every rank holds a handle created from the
    same rate matrix).  Per outer block: all-reduce of the block rows' outside mass (for the subtraction-free
    pivots), panel factorisation on the owning rank, NCCL broadcast of the factored panel, trailing update of the
    own columns on every rank -- all enqueued by the C++ engine on its streams without a host synchronisation
    between blocks.  Afterwards handle.phase_two() / rates() work as usual on every rank.
    Returns the device time in ms."""
    comm = comm or communicator(group)
    return handle.dist_phase_one(comm)


class _DeviceBlock(object):
    """exposes a raw device pointer through __cuda_array_interface__"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 3}


def _dev_tensor(ptr, n):
    import torch
    return torch.as_tensor(_DeviceBlock(ptr, n), device="cuda")


def dense_phase_one_emulated(handles):
    """The same schedule through the staged C ABI (ngt_dist_rowsum / factor / panel / apply) with the ranks emulated by
    `handles` inside ONE process on ONE GPU (tests): the collectives become explicit sums / copies between the
    handles' device buffers."""
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
