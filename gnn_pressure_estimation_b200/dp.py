"""Data-parallel plumbing: one process per GPU, snapshots sharded across ranks,
one gradient all-reduce per step.

The reference is single-process (/root/reference/gnn_pressure_estimation/train.py:306-324);
this layer is additive and its contract is "N-rank step == 1-rank step on the
concatenated batch".  Snapshots are independent (block-diagonal batch graph), so
inference needs no communication at all and training needs exactly one
collective: the SUM of the flat gradient buffer, divided by the world size
inside the Adam kernel (every snapshot masks the same number of nodes, so the
mean of the shard-mean losses is the global mean loss).

Only host logic lives here (works with gloo on CPU for tests and NCCL on GPUs).
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def shard_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous snapshot range [lo, hi) of `rank`; shards must be equal-sized so
    that averaging shard gradients equals the full-batch gradient."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} must be divisible by the world size {world} "
                         "(equal shards keep mean-of-means == global mean)")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def shard_bounds_uneven(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of `rank` when `total` need not divide by the world size (evaluation: the epoch sums
    are weighted by graphs per batch and divided by the dataset length after the all-reduce, so unequal — even empty
    — shards give the right result).  The first `total % world` ranks hold one snapshot more."""
    per, rem = divmod(int(total), int(world))
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


def bucket_ranges(num_blocks: int, buckets: int) -> List[Tuple[int, int]]:
    """Split the backward walk over blocks num_blocks-1 .. 0 into `buckets` contiguous descending ranges
    [(k_hi, k_lo), ...] of near-equal length (earlier = later blocks; the first range also carries the decoder,
    the last one the encoder).  A model without blocks has the single range (-1, 0)."""
    if num_blocks <= 0:
        return [(-1, 0)]
    buckets = max(1, min(int(buckets), num_blocks))
    out, hi = [], num_blocks - 1
    for i in range(buckets):
        n = (num_blocks - (num_blocks - 1 - hi)) // (buckets - i)    # blocks left / ranges left
        out.append((hi, hi - n + 1))
        hi -= n
    return out


def bucket_slice(num_blocks: int, param_count: int, k_hi: int, k_lo: int, offset_of_block) -> Tuple[int, int]:
    """[lo, hi) of the flat gradient buffer that is final once blocks k_hi..k_lo have been back-propagated
    (layout lin0 | block 0 | ... | block nb-1 | lin1; `offset_of_block(k)` = start of block k, k == nb -> lin1)."""
    head = num_blocks <= 0 or k_hi == num_blocks - 1
    tail = num_blocks <= 0 or k_lo == 0
    lo = 0 if tail else offset_of_block(k_lo)
    hi = param_count if head else offset_of_block(k_hi + 1)
    return lo, hi


class PeerGradients:
    """Symmetric (peer-mapped) gradient buffers for the all-reduce that is fused into the Adam kernel
    (`gatres_adam_step_peer`): two flat fp32 gradient buffers (step parity) and one uint32 flag array per rank,
    allocated with torch's CUDA symmetric memory so that every rank holds device pointers to every peer's copy
    (NVLink / NVSwitch loads).  Raises if symmetric memory is unavailable; the caller then keeps the NCCL path."""

    def __init__(self, param_count: int, group, device):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        n = (param_count + 3) // 4 * 4
        self.grads = [symm.empty(n, dtype=torch.float32, device=device) for _ in range(2)]
        self.flags = symm.empty(64, dtype=torch.int32, device=device)
        for t in self.grads + [self.flags]:
            t.zero_()
        torch.cuda.synchronize(device)
        handles = [symm.rendezvous(t, group) for t in self.grads + [self.flags]]
        self._handles = handles                                   # keep the mappings alive
        arr = C.c_void_p * self.world
        self.grad_tables = [arr(*[int(p) for p in h.buffer_ptrs]) for h in handles[:2]]
        self.flag_table = arr(*[int(p) for p in handles[2].buffer_ptrs])
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        dist.barrier(group)                                       # every rank's buffers are zeroed and mapped
        torch.cuda.synchronize(device)


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """torchrun-style rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).  -> (rank, world, local_rank)"""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local


def broadcast_parameters_(flat: Tensor, group=None, src: int = 0) -> None:
    """Make every rank start from rank `src`'s weights (flat buffer, in place)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src=src, group=group)


def allreduce_gradients_(flat_grads: Tensor, group=None, average: bool = True) -> Tensor:
    """SUM the flat gradient buffer over ranks (optionally divide by world size)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat_grads.div_(dist.get_world_size(group))
    return flat_grads


def replicas_in_sync(flat: Tensor, group=None, tol: float = 0.0) -> bool:
    """Failure detection: True iff all ranks hold the same parameters (max-min <= tol)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return True
    hi, lo = flat.clone(), flat.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    return bool((hi - lo).abs().max() <= tol)
