"""Multi-GPU plumbing: one process per GPU, image batches sharded across ranks, one all-reduce.

The forward pass has no cross-image dependency (the global pools of
/root/reference/model_library.py:409-411,421 are per image), so ranks enhance contiguous
slices of every batch independently and the only exchange step is a single SUM all-reduce of
the additive metric totals (``data_utils.reduce_metric_sums``) - NCCL over NVLink on GPUs,
gloo in the CPU tests.  Equal-size batches (the reference uses drop_remainder=True,
data_utils.py:392) make mean-over-images equal to the reference's mean of batch means.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank).

    Single-process runs (no RANK in the environment) return (0, 1, 0) without initialising.
    """
    if "RANK" not in os.environ or int(os.environ.get("WORLD_SIZE", "1")) == 1:
        return 0, 1, int(os.environ.get("LOCAL_RANK", "0"))
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        if backend == "nccl":
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_range(n, rank_, world_):
    """Contiguous slice [lo, hi) of n images owned by ``rank_``; sizes differ by at most one."""
    base, rem = divmod(n, world_)
    lo = rank_ * base + min(rank_, rem)
    return lo, lo + base + (1 if rank_ < rem else 0)


def shard_batch(tensors, rank_=None, world_=None):
    """Slice every tensor of a batch along dim 0 for this rank."""
    r = rank() if rank_ is None else rank_
    w = world() if world_ is None else world_
    n = tensors[0].shape[0]
    lo, hi = shard_range(n, r, w)
    return [t[lo:hi] for t in tensors]


def all_reduce_totals(totals):
    """SUM all-reduce of the fp64 totals vector, in place, stream-ordered (no host sync)."""
    if world() > 1:
        dist.all_reduce(totals, op=dist.ReduceOp.SUM)
    return totals


def barrier():
    if world() > 1:
        dist.barrier()


# ------------------------------------------------------------------ spatial sharding of one large image
SPATIAL_HALO = 72      # receptive-field radius of the coefficient path (~70 px) rounded up to the network stride


def spatial_shards(height, world_, stride=8, halo=SPATIAL_HALO):
    """Row bands of ONE image for ``world_`` ranks: a list of dicts
    ``own=(a, b)`` rows of the image the rank owns, ``slab=(s0, s1)`` rows it must read (own + halo, clipped to the
    image), ``own_in_slab=(a - s0, b - s0)``.  All numbers are multiples of ``stride``."""
    assert height % stride == 0 and halo % stride == 0
    units = height // stride
    base, rem = divmod(units, world_)
    out, a = [], 0
    for r in range(world_):
        b = a + (base + (1 if r < rem else 0)) * stride
        s0, s1 = max(0, a - halo), min(height, b + halo)
        out.append(dict(own=(a, b), slab=(s0, s1), own_in_slab=(a - s0, b - s0)))
        a = b
    return out


def all_reduce_sum(t):
    """SUM all-reduce of a CUDA tensor, in place, stream-ordered; identity for a single process."""
    if world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t
