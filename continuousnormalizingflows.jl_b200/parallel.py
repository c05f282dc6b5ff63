"""Batch sharding for multi-GPU runs: one process per GPU, samples split by
columns, parameters replicated.

The path has exactly one exchange step: the training gradient (and the scalar
loss) are summed over ranks.  Each shard is told its global column offset (so the
in-kernel Philox noise is the one the unsharded batch would draw) and the global
batch size (so the shards' losses and gradients SUM to the unsharded result --
exactly for fixed-step solves; with adaptive stepping every rank's controller sees
its own shard's error norm unless the communicator of ``icnf_group_*`` is attached,
so the shards then agree with the unsharded solve to solver tolerance only);
inference and generate need no communication at all.  STEER draws one t1 per solve
for the whole batch (base_icnf.jl:23-43): seed ``ICNF(rng=...)`` identically on every
rank so that all shards integrate to the same t1.  The collective is
``torch.distributed.all_reduce`` -- NCCL over NVLink for CUDA tensors, gloo for
host arrays.
"""

from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None

from . import api


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Columns [lo, hi) of an n-column batch owned by `rank`: contiguous, sizes differ by at most 1."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return (n * rank) // world, (n * (rank + 1)) // world


def all_reduce_sum(x, group=None):
    """In-place sum over ranks of a torch tensor (CUDA -> NCCL, CPU -> gloo) or a numpy array (through gloo)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x
    if isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        x[...] = t.numpy()
        return x
    dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
    return x


def group_join(icnf, rank: int, world: int, group=None):
    """Attach the library's own communicator to ``icnf`` (``icnf_group_join``): rank 0 draws the id, torch.distributed
    (whatever backend is initialised) only ships its 128 bytes.  After this, ``dp_loss_and_gradient`` sums
    [gradient; loss] inside the library -- in the gradient-reduction kernel over NVLink peer memory for small
    networks, with ncclAllReduce on the same stream for large ones -- and torch.distributed is out of the step."""
    if world == 1:
        return
    uid = [api.group_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0, group=group)
    api.group_join_id(icnf, uid[0], rank, world)


def dp_loss_and_gradient(icnf, mode, xs_local, *args, rank: int, world: int, global_batch: int,
                         local_fn: Optional[Callable] = None, group=None, **kw):
    """Data-parallel training step: local loss/gradient on this rank's columns, then
    ONE all-reduce of [gradient; loss].  ``xs_local`` holds columns
    ``shard_bounds(global_batch, rank, world)`` of the global batch.  ``local_fn``
    defaults to ``api.loss_and_gradient`` (the CUDA path); tests inject a CPU stand-in
    to exercise the protocol under gloo.  A handle that joined a group (``group_join``) does
    the exchange inside the library; otherwise ``torch.distributed.all_reduce`` is used."""
    lo, hi = shard_bounds(global_batch, rank, world)
    if xs_local.shape[1] != hi - lo:
        raise ValueError(f"rank {rank} expects {hi - lo} columns, got {xs_local.shape[1]}")
    if local_fn is None and getattr(icnf, "_group", None) == (rank, world) and world > 1:
        return api.loss_and_gradient(icnf, mode, xs_local, *args, sample_offset=lo, global_batch=global_batch,
                                     data_parallel=True, **kw)
    fn = local_fn or api.loss_and_gradient
    l, g = fn(icnf, mode, xs_local, *args, sample_offset=lo, global_batch=global_batch, **kw)
    if torch is not None and isinstance(g, torch.Tensor):
        packed = getattr(icnf, "_grad_loss_buf", None)
        if packed is None or packed.data_ptr() != g.data_ptr():     # not the shared [dtheta; loss] buffer
            packed = torch.cat([g.reshape(-1), l.reshape(1)])
        all_reduce_sum(packed, group)
        return packed[-1], packed[:-1]
    packed = np.concatenate([np.asarray(g, dtype=np.float32).reshape(-1), np.asarray([l], dtype=np.float32)])
    all_reduce_sum(packed, group)
    return float(packed[-1]), packed[:-1]
