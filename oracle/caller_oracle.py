"""CPU oracle for the caller-side pieces next to the hot path: masks, (de)normalisation and the seven
training / evaluation metrics.  TEST INFRASTRUCTURE ONLY (same import rule as gatres_oracle.py).

**PARITY PINNED** for this file: the reference module it restates
(/root/reference/gnn_pressure_estimation/utils/auxil.py) is plain numpy / PyTorch and imports in this
container, so `tests/golden/make_golden_caller.py` runs the REAL reference functions and commits their
outputs (`tests/golden/caller_ref.npz`); `tests/test_caller.py` holds this restatement to those vectors.

Reference symbols restated (all in utils/auxil.py):
  scale / descale                     :18-64
  calculate_nse                       :101-107
  calculate_rmse                      :110-111
  calculate_rel_error                 :114-118
  calculate_accuracy                  :121-124
  calculate_correlation_coefficient   :127-135
  calculate_r2                        :138-140
  mask_nodes / generate_batch_mask    :143-182
  get_metric_fn_collection            :185-203  (the order / names of the seven metrics)
and the way train.py:177-198 / evaluation.py:326-338 apply them: on the DESCALED predictions and targets
of the masked nodes of one batch.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch
from torch import Tensor

METRIC_NAMES = ("error", "0.1", "corr", "r2", "mae", "rmse", "mynse")      # auxil.py:194-202, in this order


def scale(data, norm_type: str = "minmax", mean=None, std=None, min=None, max=None, eps: float = 1e-8):
    """auxil.py:18-39"""
    assert norm_type in ("minmax", "znorm")
    if norm_type == "minmax":
        return (data - min) / (max - min)
    return (data - mean) / (std + eps)


def descale(scaled, norm_type: Optional[str] = "minmax", mean=None, std=None, min=None, max=None):
    """auxil.py:42-64 (any other norm_type is the identity)"""
    if norm_type == "minmax":
        return scaled * (max - min) + min
    if norm_type == "znorm":
        return scaled * std + mean
    return scaled


def metrics(y_pred: Tensor, y_true: Tensor, threshold: float = 0.1) -> Dict[str, Tensor]:
    """The seven metrics of get_metric_fn_collection on one batch of (descaled) masked predictions."""
    p, t = y_pred, y_true
    err = (t - p).abs()
    keep = t.abs() > 0.01                                                  # :116
    rel = (err[keep] / t[keep]).abs().mean()                                # :117-118 (nan when nothing is kept)
    acc = (err <= t * threshold).to(p.dtype).mean()                         # :122-124 (negative targets never count)
    vx, vy = p - p.mean(), t - t.mean()                                     # :128-129
    corr = torch.clamp((vx * vy).sum() / (torch.sqrt((vx ** 2).sum()) * torch.sqrt((vy ** 2).sum())), -1.0, 1.0)
    mae = err.mean()                                                        # F.l1_loss
    rmse = torch.sqrt(((p - t) ** 2).mean())                                # :111
    pr, tr = p.reshape(-1), t.reshape(-1)
    nse = 1.0 - ((pr - tr) ** 2).sum() / (((tr - tr.mean()) ** 2).sum() + 1e-12)   # :104-107
    return {"error": rel, "0.1": acc, "corr": corr, "r2": corr ** 2, "mae": mae, "rmse": rmse, "mynse": nse}


def mask_nodes(num_nodes: int, masking_rate: float, required_idx: Sequence[int] = ()) -> np.ndarray:
    """auxil.py:143-163; consumes the GLOBAL numpy RNG exactly like the reference."""
    required_idx = list(required_idx)
    mask_length = int(num_nodes * masking_rate) - len(required_idx)
    assert mask_length > 0
    selected = [i for i in range(num_nodes) if i not in set(required_idx)]   # list(set(range(n)) - required): ascending
    idx = np.random.choice(selected, mask_length, replace=False)
    mask = np.zeros(num_nodes)
    mask[idx] = 1
    mask[required_idx] = 1
    assert int(mask.sum()) == int(num_nodes * masking_rate)
    return mask.astype(bool)


def generate_batch_mask(num_nodes: Sequence[int], mask_rate: float, required_idx: Sequence[int] = ()) -> np.ndarray:
    """auxil.py:166-182: one exact-count mask per snapshot, concatenated."""
    return np.hstack([mask_nodes(int(n), mask_rate, required_idx) for n in num_nodes])


# ---- checker for the DEVICE mask generator (our design, not a reference function) --------------------------
# gnn_pressure_estimation_b200/csrc/caller.cu draws, for node (b, i), the key splitmix64(seed, step, b*N + i) >> 32
# and selects the `count` smallest keys of each snapshot, ties by node index.  Restated here with numpy uint64
# arithmetic so the CUDA kernel can be checked bit for bit; the reference contract it must satisfy (exact count
# per snapshot, uniform without replacement - auxil.py:143-163) is tested through properties.
_U64 = np.uint64


def device_mask_keys(seed: int, step: int, rows: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (_U64(seed) ^ (_U64(step) * _U64(0xD1B54A32D192ED03))) + (rows.astype(np.uint64) + _U64(1)) * _U64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> _U64(30))) * _U64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> _U64(27))) * _U64(0x94D049BB133111EB)
        z = z ^ (z >> _U64(31))
    return np.maximum((z >> _U64(32)).astype(np.uint32), np.uint32(1))      # key 0 is reserved for required nodes


def device_mask_reference(seed: int, step: int, batch: int, num_nodes: int, count: int,
                          required_idx: Sequence[int] = ()) -> np.ndarray:
    keys = device_mask_keys(seed, step, np.arange(batch * num_nodes)).reshape(batch, num_nodes)
    keys[:, list(required_idx)] = 0
    order = np.argsort(keys, axis=1, kind="stable")                  # ties keep node order
    mask = np.zeros((batch, num_nodes), dtype=bool)
    np.put_along_axis(mask, order[:, :count], True, axis=1)
    return mask.reshape(-1)


def test_one_epoch_oracle(model, snapshots: Tensor, edge_index: Tensor, batch_size: int, mask_rate: float,
                          norm_type=None, mean=None, std=None, min_val=None, max_val=None,
                          required_idx: Sequence[int] = (), use_same_mask: bool = False):
    """evaluation.py:300-341 on a [S, N] snapshot set with a CPU model (the GATRes oracle): per batch a mask from the
    global numpy RNG, x[mask] = 0, forward, MSE and the seven metrics on the (descaled) masked nodes, every per-batch
    value weighted by num_graphs and divided by the dataset length.  -> (loss, {metric: value})"""
    S, N = snapshots.shape
    total_loss, totals = 0.0, {k: 0.0 for k in METRIC_NAMES}
    all_mask = None
    with torch.no_grad():
        for s0 in range(0, S, batch_size):
            y = snapshots[s0:s0 + batch_size].reshape(-1, 1)
            B = y.numel() // N
            if all_mask is None or not use_same_mask:
                all_mask = torch.from_numpy(generate_batch_mask([N] * B, mask_rate, required_idx))
            mask = all_mask[:B * N]
            x1 = y.clone()
            x1[mask] = 0
            ei = torch.cat([edge_index + b * N for b in range(B)], dim=1)
            out = model(x1, ei, None, None)
            y_pred, y_true = out[mask], y[mask]
            kw = dict(norm_type=norm_type, mean=mean, std=std, min=min_val, max=max_val)
            m = metrics(descale(y_pred, **kw), descale(y_true, **kw))
            total_loss += float(torch.nn.functional.mse_loss(y_pred, y_true)) * B
            for k in METRIC_NAMES:
                totals[k] += float(m[k]) * B
    return total_loss / S, {k: v / S for k, v in totals.items()}
