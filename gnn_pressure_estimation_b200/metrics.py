"""Device-side masks and metrics of the reference loop.

Host mirror of /root/reference/gnn_pressure_estimation/utils/auxil.py for the pieces that sit next to the hot
path in `train_one_epoch` / `test_one_epoch`:

  generate_batch_mask  auxil.py:166-182 (per batch at train.py:171-172, evaluation.py:312-322)
                       -> `generate_batch_mask(...)` below: same contract (exactly int(N*rate) nodes per snapshot,
                       uniform without replacement), drawn on the device by `gatres_generate_mask`.
  descale              auxil.py:42-64 -> `descale_affine(norm_type, ...)`: every norm is v*scale + shift.
  get_metric_fn_collection / calculate_*   auxil.py:101-140,185-203
                       -> `MaskedMetrics`: one fused two-pass reduction (`gatres_masked_metrics`) instead of
                       ~40 micro-kernels and 7 host syncs per batch; same names, same order.

CUDA only: there is no CPU path (the oracle in oracle/caller_oracle.py is test infrastructure).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from . import _lib
from ._lib import call, ptr, stream

METRIC_NAMES = ("error", "0.1", "corr", "r2", "mae", "rmse", "mynse")     # auxil.py:194-202


def descale_affine(norm_type: Optional[str], mean=None, std=None, min=None, max=None) -> Tuple[float, float]:
    """(scale, shift) with descale(v) == v * scale + shift  (auxil.py:42-64; unknown norm = identity)."""
    if norm_type == "minmax":
        if min is None or max is None:
            raise ValueError("min and max values are missing")
        return float(max) - float(min), float(min)
    if norm_type == "znorm":
        if mean is None or std is None:
            raise ValueError("mean and std values are missing")
        return float(std), float(mean)
    return 1.0, 0.0


def mask_count(num_nodes: int, mask_rate: float) -> int:
    """masked nodes per snapshot, auxil.py:154 with required_idx=[]"""
    return int(num_nodes * mask_rate)


def required_flags(num_nodes: int, required_idx: Sequence[int], device) -> Optional[Tensor]:
    """uint8 [num_nodes] flags of the nodes every mask must contain (`required_idx`, auxil.py:143-161)"""
    if len(required_idx) == 0:
        return None
    f = torch.zeros(num_nodes, dtype=torch.uint8)
    f[torch.as_tensor(list(required_idx), dtype=torch.long)] = 1
    return f.to(device)


def generate_batch_mask(batch: int, num_nodes: int, mask_rate: float, seed: int, step: int = 0,
                        out: Optional[Tensor] = None, step_dev: Optional[Tensor] = None,
                        required: Optional[Tensor] = None, device=None) -> Tensor:
    """uint8 [batch*num_nodes]: exactly int(num_nodes*mask_rate) ones per snapshot, chosen on the device;
    `required` = required_flags(...) of the nodes that are always masked."""
    count = mask_count(num_nodes, mask_rate)
    n_req = int(required.sum()) if required is not None else 0
    if count - n_req <= 0:
        raise ValueError("mask_rate leaves no node to draw (the reference asserts mask_length > 0)")
    if out is None:
        out = torch.empty(batch * num_nodes, dtype=torch.uint8, device=device if device is not None else "cuda")
    call("gatres_generate_mask", seed & (2 ** 64 - 1), step & (2 ** 64 - 1), ptr(step_dev), ptr(required), batch,
         num_nodes, count, ptr(out), stream())
    return out


def numpy_batch_mask(num_nodes: Sequence[int], mask_rate: float, required_idx: Sequence[int] = ()) -> np.ndarray:
    """NumPy-compatible mode: the reference's own host procedure (auxil.py:143-182) on the global NumPy RNG,
    for runs that must reproduce the reference's mask stream draw for draw."""
    req = list(required_idx)
    parts = []
    for n in num_nodes:
        n = int(n)
        length = int(n * mask_rate) - len(req)
        assert length > 0
        candidates = [i for i in range(n) if i not in set(req)]
        m = np.zeros(n, dtype=bool)
        m[np.random.choice(candidates, length, replace=False)] = True
        m[req] = True
        parts.append(m)
    return np.hstack(parts)


class MaskedMetrics:
    """The seven reference metrics over the masked nodes of a batch, computed on the device in one call.

    `update(out, y, mask)` enqueues the reduction and returns the device tensor
    [error, 0.1, corr, r2, mae, rmse, mynse, count]; `as_dict()` reads it back (one sync) with the reference's
    key names (`{prefix}_error`, ...)."""

    def __init__(self, device, norm_type: Optional[str] = "znorm", mean=None, std=None, min=None, max=None,
                 threshold: float = 0.1, prefix: str = "tr"):
        self.scale, self.shift = descale_affine(norm_type, mean, std, min, max)
        self.threshold, self.prefix = float(threshold), prefix
        lib = _lib.load()
        self._scratch = torch.empty(int(lib.gatres_metrics_scratch_doubles()), dtype=torch.float64, device=device)
        self.values = torch.zeros(8, dtype=torch.float32, device=device)

    def update(self, out: Tensor, y: Tensor, mask: Optional[Tensor]) -> Tensor:
        out, y = out.reshape(-1), y.reshape(-1)
        if mask is not None:
            mask = mask.reshape(-1)
            mask = mask.view(torch.uint8) if mask.dtype == torch.bool else mask
        call("gatres_masked_metrics", ptr(out), ptr(y), ptr(mask), out.numel(), self.scale, self.shift, self.threshold,
             ptr(self._scratch), ptr(self.values), stream())
        return self.values

    def as_dict(self) -> Dict[str, float]:
        v = self.values.cpu().tolist()
        return {f"{self.prefix}_{k}": v[i] for i, k in enumerate(METRIC_NAMES)}
