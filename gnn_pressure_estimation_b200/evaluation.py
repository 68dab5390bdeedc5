"""Evaluation path on the B200 kernels.

Host mirror of `test_one_epoch` (/root/reference/gnn_pressure_estimation/evaluation.py:240-351) and of its
`Timer` (utils/timer.py:12-66) for a snapshot set that is RESIDENT on the device as one [S, N] tensor instead of
S PyG `Data` objects behind a DataLoader: per batch the reference clones x, draws a mask on the host, zeroes the
masked inputs, runs the timed forward, and computes MSE + seven metrics on the descaled masked nodes with ~50
micro-kernels and 8 host syncs.  Here a batch is: (device or NumPy-compatible) mask -> `gatres_apply_mask` ->
the model's forward kernels -> `gatres_masked_mse` + `gatres_masked_metrics`; the per-batch scalars stay on the
device until the epoch ends.  With a process group the snapshot set is sharded across ranks (contiguous
shards, no communication during the forward passes) and the epoch sums are all-reduced once at the end.

Aggregation is the reference's: every per-batch value is weighted by the batch's number of graphs and divided by
the dataset length (evaluation.py:333-341), so a smaller last batch is handled the same way.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist
from torch import Tensor

from . import dp as _dp
from . import metrics as _metrics
from ._lib import call, ptr, stream


class Timer:
    """utils/timer.py:12-66: CUDA-event latency of the wrapped inference call, warm-ups before the first one."""

    def __init__(self) -> None:
        self.reset()

    def reset(self) -> None:
        self.starter, self.ender = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.timings: List[float] = []
        self.num_graphs: List[int] = []
        self.finished_warmup = False

    def auto_measure(self, inference_func: Callable, num_graphs_per_batch: int, gpu_warmup_times: int = 10) -> Callable:
        def inference(*args, **kwargs):
            if gpu_warmup_times > 0 and not self.finished_warmup:
                for _ in range(gpu_warmup_times):
                    inference_func(*args, **kwargs)
                self.finished_warmup = True
            self.starter.record()
            results = inference_func(*args, **kwargs)
            self.ender.record()
            torch.cuda.synchronize()
            self.timings.append(self.starter.elapsed_time(self.ender))
            self.num_graphs.append(num_graphs_per_batch)
            return results
        return inference

    def compute_time(self, len_dataset: int) -> float:
        """graph-weighted mean BATCH latency in ms (timer.py:43-51)"""
        return compute_time(self.timings, self.num_graphs, len_dataset)

    def compute_throughput(self, len_dataset: int) -> float:
        """the reference's figure (timer.py:53-66): len(timings) * max(num_graphs) / weighted mean latency —
        equal to snapshots/s only when the epoch is a single batch; kept for drop-in logs"""
        return compute_throughput(self.timings, self.num_graphs, len_dataset)

    def snapshots_per_second(self) -> float:
        """true throughput: snapshots processed / time spent in the timed forward calls"""
        return float(sum(self.num_graphs)) / (float(sum(self.timings)) / 1000.0)


def compute_time(timings: Sequence[float], num_graphs: Sequence[int], len_dataset: int) -> float:
    assert len(timings) == len(num_graphs) and len_dataset > 0
    return float(np.array(timings).dot(np.array(num_graphs)) / len_dataset)


def compute_throughput(timings: Sequence[float], num_graphs: Sequence[int], len_dataset: int) -> float:
    assert len(timings) == len(num_graphs) and len_dataset > 0
    total = float(np.sum(np.array(timings) * np.array(num_graphs) / len_dataset / 1000))
    return float(len(timings) * max(num_graphs)) / total


def test_one_epoch(model, snapshots: Tensor, edge_index: Tensor, batch_size: int, mask_rate: float,
                   norm_type: Optional[str] = "znorm", mean=None, std=None, min_val=None, max_val=None,
                   required_idx: Sequence[int] = (), prefix: str = "test", gpu_warmup_times: int = 10,
                   use_same_mask: bool = False, mask_source: str = "numpy", seed: int = 0,
                   process_group=None) -> Tuple[float, Dict[str, float]]:
    """One evaluation trial over `snapshots` ([S, N] normalised pressures on the device; x = y as in
    utils/auxil.py:96-97).  Returns (loss, metrics) with the reference's keys `{prefix}_error ... {prefix}_mynse,
    {prefix}_time, {prefix}_throughput` (+ `_sensor` postfix when required_idx is given, evaluation.py:288-291,343)
    plus `{prefix}_snapshots_per_s`.  mask_source: "numpy" = the reference's host procedure on the global NumPy RNG;
    "device" = gatres_generate_mask keyed by (seed, batch index)."""
    if not snapshots.is_cuda:
        raise RuntimeError("the snapshot set must be resident on the GPU (no CPU path exists)")
    if mask_source not in ("numpy", "device"):
        raise ValueError("mask_source must be 'numpy' or 'device'")
    dev = snapshots.device
    S_total, N = snapshots.shape
    rank, world = (dist.get_rank(process_group), dist.get_world_size(process_group)) if process_group is not None else (0, 1)
    lo, hi = _dp.shard_bounds_uneven(S_total, rank, world)     # a validation split need not divide by the world size
    shard = snapshots[lo:hi].contiguous()
    S = hi - lo
    count = _metrics.mask_count(N, mask_rate)
    required = _metrics.required_flags(N, required_idx, dev)
    postfix = "_sensor" if len(required_idx) else ""
    model.eval()
    ei_dev = edge_index.to(dev)
    mm = _metrics.MaskedMetrics(dev, norm_type, mean=mean, std=std, min=min_val, max=max_val, prefix=prefix)
    timer = Timer()
    totals = torch.zeros(9, dtype=torch.float64, device=dev)        # loss, 7 metrics, (unused) — each x num_graphs
    loss_buf, loss_part = torch.zeros(1, device=dev), torch.empty(1024, device=dev)
    all_mask: Optional[Tensor] = None
    collated: Dict[int, Tensor] = {}

    with torch.no_grad():
        for bi, s0 in enumerate(range(0, S, batch_size)):
            y = shard[s0:s0 + batch_size].reshape(-1)
            Bb = y.numel() // N
            if Bb not in collated:                                    # PyG collation of the template
                off = (torch.arange(Bb, device=dev) * N).repeat_interleave(ei_dev.size(1))
                collated[Bb] = ei_dev.repeat(1, Bb) + off
            if all_mask is None or not use_same_mask:
                if mask_source == "numpy":
                    all_mask = torch.from_numpy(_metrics.numpy_batch_mask([N] * Bb, mask_rate, required_idx)).view(torch.uint8).to(dev)
                else:
                    all_mask = _metrics.generate_batch_mask(Bb, N, mask_rate, seed + rank, bi, required=required, device=dev)
            mask = all_mask[:Bb * N]
            x1 = torch.empty_like(y)
            call("gatres_apply_mask", ptr(y), ptr(mask), ptr(x1), y.numel(), stream())          # evaluation.py:312-322
            wrapped = timer.auto_measure(model, num_graphs_per_batch=Bb, gpu_warmup_times=gpu_warmup_times)
            out = wrapped(x1.view(-1, 1), collated[Bb], None, None).reshape(-1)
            d_out = torch.empty_like(out)
            call("gatres_masked_mse", ptr(out), ptr(y), ptr(mask), out.numel(), Bb * count, ptr(d_out), ptr(loss_buf),
                 ptr(loss_part), stream())                                                   # criterion on masked nodes
            vals = mm.update(out, y, mask)
            totals[0] += loss_buf[0].double() * Bb
            totals[1:8] += vals[:7].double() * Bb

    if world > 1:
        dist.all_reduce(totals, group=process_group)
    t = (totals / S_total).cpu().tolist()
    result = {f"{prefix}_{k}": t[1 + i] for i, k in enumerate(_metrics.METRIC_NAMES)}
    if S > 0:
        time_ms, thr, sps = timer.compute_time(S), timer.compute_throughput(S), timer.snapshots_per_second()
    else:                                                             # empty shard (fewer snapshots than ranks)
        time_ms, thr, sps = 0.0, float("inf"), float("inf")
    if world > 1:
        agg = torch.tensor([time_ms, -thr, -sps], dtype=torch.float64, device=dev)
        dist.all_reduce(agg, op=dist.ReduceOp.MAX, group=process_group)                      # slowest rank
        time_ms, thr, sps = float(agg[0]), -float(agg[1]) * world, -float(agg[2]) * world
    result[f"{prefix}_time"] = time_ms
    result[f"{prefix}_throughput"] = thr
    result[f"{prefix}_snapshots_per_s"] = sps
    return t[0], {k + postfix: v for k, v in result.items()}
