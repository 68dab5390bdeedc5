"""Training driver on the B200 kernels: the reference's experiment loop for GATRes on a device-resident snapshot set.

Host mirror of `train_one_epoch` / `internal_train` (/root/reference/gnn_pressure_estimation/train.py:112-202,
282-532) for the pieces that surround the hot path: epoch loop over shuffled batches, per-batch mask, one optimizer
step, loss and the seven metrics weighted by graphs per batch and divided by the dataset length (train.py:190-202),
validation with `test_one_epoch`, best / last checkpoints in the reference's format (`save_checkpoint(path, **kwargs)`,
utils/auxil.py:223-233; keys as written at train.py:433-451), early stopping (utils/early_stopping.py:31-78).
The reference script itself cannot run in this image (it imports torch_geometric, zarr and wntr at module top); with
those installed its `train.py` keeps working on the drop-in `GraphModels.GATResMeanConv`.  This driver is the additive
fast path: the snapshot set lives on the GPU (`snapshot_store.SnapshotSet`), a batch is a slice of it, the step is
one captured CUDA graph (`train_step.TrainStep`), nothing is synchronised until the epoch ends.

    python -m gnn_pressure_estimation_b200.train --model gatres_small --synthetic 4096 --epochs 20
    python -m gnn_pressure_estimation_b200.train --model gatres_small --input_path net.inp --dataset_path net.zip
    torchrun --nproc-per-node 8 -m gnn_pressure_estimation_b200.train ...        # snapshots sharded over ranks
"""
from __future__ import annotations

import argparse
import math
import os
import time
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist
from torch import Tensor

from . import ConfigModels, dp as _dp, evaluation as _evaluation, metrics as _metrics, topology as _topology
from .snapshot_store import SnapshotSet
from .train_step import TrainStep


class EarlyStopping:
    """utils/early_stopping.py:31-78 (mode 'min', absolute min_delta); patience 0 disables it."""

    def __init__(self, min_delta: float = 0.0, patience: int = 10):
        self.min_delta, self.patience = min_delta, patience
        self.best: Optional[float] = None
        self.num_bad_epochs = 0

    def step(self, metric: float) -> bool:
        if self.patience == 0:
            return False
        if self.best is None:
            self.best = metric
            return False
        if math.isnan(metric):
            return True
        if metric < self.best - self.min_delta:
            self.num_bad_epochs, self.best = 0, metric
        else:
            self.num_bad_epochs += 1
        return self.num_bad_epochs >= self.patience


def adam_state_dict(step: TrainStep) -> dict:
    """the flat Adam buffers in torch.optim.Adam's state_dict layout (what train.py:436 stores).

    torch.optim.Adam(model.parameters()) numbers its state by position in ``model.parameters()`` (a module's own
    parameters before its children's: att_src, att_dst, bias, lin_src.weight per GATConv), which is NOT the order of
    the flat buffer (``ordered_parameters()``: weight first) — so every parameter is looked up by its offset."""
    offsets, off = {}, 0
    for p in step.model.ordered_parameters():
        offsets[id(p)] = off
        off += p.numel()
    state = {}
    t = float(step.step_count.item())
    for i, p in enumerate(step.model.parameters()):
        o, n = offsets[id(p)], p.numel()
        state[i] = {"step": torch.tensor(t), "exp_avg": step.exp_avg[o:o + n].view_as(p).clone(),
                    "exp_avg_sq": step.exp_avg_sq[o:o + n].view_as(p).clone()}
    group = {"lr": step.lr, "betas": tuple(step.betas), "eps": step.eps, "weight_decay": step.wd, "amsgrad": False,
             "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
             "params": list(range(len(state)))}
    return {"state": state, "param_groups": [group]}


def save_checkpoint(path: str, **kwargs) -> str:
    """utils/auxil.py:223-233"""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(kwargs, path)
    return path


def load_checkpoint(path: str, model) -> Tuple[torch.nn.Module, dict]:
    """utils/auxil.py:206-220 (weights_only=False: the reference's checkpoints hold numpy scalars)"""
    assert path[-4:] == ".pth" and model is not None
    cp = torch.load(path, map_location="cpu", weights_only=False)
    model.load_state_dict(cp["model_state_dict"])
    return model, cp


def train_one_epoch(steps: Dict[int, TrainStep], snapshots: Tensor, batch_size: int, mask_rate: float,
                    generator: Optional[torch.Generator], mask_source: str = "device", drop_last: bool = False,
                    prefix: str = "tr") -> Tuple[float, Dict[str, float]]:
    """train.py:159-202 on the device-resident set `snapshots` [S, N].  `steps` maps a batch size to its TrainStep (the
    full batch and, unless drop_last, the smaller last batch).  -> (loss, metrics), both divided by the dataset length."""
    dev = snapshots.device
    S, N = snapshots.shape
    order = torch.randperm(S, generator=generator).to(dev) if generator is not None else torch.arange(S, device=dev)
    totals = torch.zeros(8, dtype=torch.float64, device=dev)            # loss, 7 metrics — each x num_graphs
    seen = 0
    for s0 in range(0, S, batch_size):
        idx = order[s0:s0 + batch_size]
        B = int(idx.numel())
        if B not in steps:
            if drop_last:
                continue
            raise KeyError(f"no TrainStep for a batch of {B} snapshots")
        ts = steps[B]
        y = snapshots[idx].reshape(-1)
        if mask_source == "numpy":
            mask = torch.from_numpy(_metrics.numpy_batch_mask([N] * B, mask_rate)).view(torch.uint8).to(dev)
            loss = ts.step(y, y, mask)
        else:
            loss = ts.step(y, y, None)
        totals[0] += loss[0].double() * B
        if ts.metrics is not None:
            totals[1:8] += ts.metrics.values[:7].double() * B
        seen += B
    if ts.pg is not None and ts.world > 1:
        cnt = torch.tensor([float(seen)], dtype=torch.float64, device=dev)
        dist.all_reduce(totals, group=ts.pg)
        dist.all_reduce(cnt, group=ts.pg)
        seen = int(cnt.item())
    t = (totals / max(seen, 1)).cpu().tolist()
    return t[0], {f"{prefix}_{k}": t[1 + i] for i, k in enumerate(_metrics.METRIC_NAMES)}


def smooth_synthetic_snapshots(wn: _topology.WaterNetwork, num: int, seed: int = 0, modes: int = 12, noise: float = 0.05
                               ) -> Tuple[np.ndarray, np.ndarray]:
    """Learnable stand-in for simulated pressures (no EPANET / datasets here): every snapshot is a random combination
    of the `modes` smoothest eigenvectors of the network's graph Laplacian (pressure fields vary smoothly along pipes)
    plus white noise, shifted and scaled to metres of head.  -> (snapshots [num, N] float64, edge_index)."""
    ei, names = _topology.reference_edge_index(wn, "keep_junction")
    n = len(names)
    A = np.zeros((n, n))
    A[ei[0], ei[1]] = 1.0
    L = np.diag(A.sum(1)) - A
    _, vec = np.linalg.eigh(L)
    basis = vec[:, 1:1 + modes]                                         # skip the constant mode
    rng = np.random.RandomState(seed)
    coef = rng.randn(num, modes) * np.linspace(1.0, 0.3, modes)
    data = coef @ basis.T * math.sqrt(n) + noise * rng.randn(num, n)
    return 60.0 + 8.0 * data, ei


def fit(model, train_set: SnapshotSet, valid_set: SnapshotSet, batch_size: int = 32, epochs: int = 10, lr: float = 5e-4,
        weight_decay: float = 6e-6, mask_rate: float = 0.95, patience: int = 0, save_path: Optional[str] = None,
        variant: str = "b200", mask_source: str = "device", seed: int = 0, process_group=None, log_every: int = 5,
        drop_last: bool = False) -> Dict[str, object]:
    """internal_train (train.py:282-532) for GATRes: returns the history and the best validation result."""
    dev = train_set.snapshots.device
    rank, world = (dist.get_rank(process_group), dist.get_world_size(process_group)) if process_group is not None else (0, 1)
    N = train_set.num_nodes
    model = model.to(dev)
    topo = model.set_topology(train_set.edge_index.to(dev), N)
    lo, hi = _dp.shard_bounds(len(train_set) - len(train_set) % world, rank, world)
    shard = train_set.snapshots[lo:hi]
    norm = dict(norm_type=train_set.norm_type, mean=train_set.mean, std=train_set.std, min=train_set.min, max=train_set.max)
    count = _metrics.mask_count(N, mask_rate)

    def make_step(B, share=None):
        mm = _metrics.MaskedMetrics(dev, prefix="tr", **norm)
        ts = TrainStep(model, topo, B, count, lr=lr, weight_decay=weight_decay, process_group=process_group,
                       device_mask_seed=seed + 1 if mask_source == "device" else None, metrics=mm, share_state_with=share)
        ts.capture()
        return ts

    steps = {batch_size: make_step(batch_size)}
    rem = shard.shape[0] % batch_size
    if rem and not drop_last:
        steps[rem] = make_step(rem, steps[batch_size])
    gen = torch.Generator().manual_seed(seed + 17 * rank)
    stopper = EarlyStopping(patience=patience)
    best = {"loss": float("inf"), "epoch": 0, "metrics": {}}
    history = []
    t0 = time.time()
    for epoch in range(1, epochs + 1):
        model.train()
        if world > 1:
            # rank 0 may still be writing checkpoints / printing: the peer-memory Adam kernel spins on every rank's
            # flag inside a captured graph, so ranks enter an epoch's first step together instead of one of them
            # waiting on the device for a slow filesystem
            dist.barrier(process_group)
        tr_loss, tr_metrics = train_one_epoch(steps, shard, batch_size, mask_rate, gen, mask_source, drop_last)
        val_loss, val_metrics = _evaluation.test_one_epoch(
            model, valid_set.snapshots, valid_set.edge_index, batch_size, mask_rate, prefix="val", gpu_warmup_times=0,
            mask_source=mask_source, seed=seed + 1000 + epoch, process_group=process_group,
            norm_type=norm["norm_type"], mean=norm["mean"], std=norm["std"], min_val=norm["min"], max_val=norm["max"])
        history.append({"epoch": epoch, "tr_loss": tr_loss, "val_loss": val_loss, **tr_metrics, **val_metrics})
        ckpt = dict(model_state_dict=model.state_dict(), optimizer_state_dict=adam_state_dict(steps[batch_size]),
                    mean=train_set.mean, std=train_set.std, min=train_set.min, max=train_set.max, edge_attrs=None,
                    edge_mean=None, edge_std=None, edge_min=None, edge_max=None, norm_type=train_set.norm_type)
        if val_loss < best["loss"]:
            best = {"loss": val_loss, "epoch": epoch, "metrics": val_metrics}
            if save_path and rank == 0:
                save_checkpoint(os.path.join(save_path, f"best_{model.name}_{variant}.pth"), epoch=epoch, loss=val_loss,
                                val_metric_dict=val_metrics, val_record_metric_dict={}, **ckpt)
        if epoch == 1 or epoch % log_every == 0:
            if rank == 0:
                print(f"epoch {epoch:4d}  tr_loss {tr_loss:.5f}  val_loss {val_loss:.5f}  val_mae {val_metrics['val_mae']:.4f}  "
                      f"val_r2 {val_metrics['val_r2']:.4f}  {time.time() - t0:.1f}s", flush=True)
            if save_path and rank == 0 and not math.isnan(tr_loss):
                save_checkpoint(os.path.join(save_path, f"last_{model.name}_{variant}.pth"), epoch=best["epoch"],
                                loss=best["loss"], val_metric_dict=val_metrics, val_record_metric_dict={}, **ckpt)
        if stopper.step(val_loss):
            break
    return {"history": history, "best": best, "steps": steps}


def get_arguments() -> argparse.Namespace:
    ap = argparse.ArgumentParser(description="GATRes training on the B200 kernels (mirror of the reference's train.py flags)")
    ap.add_argument("--model", default="gatres_small", choices=["gatres_small", "gatres_large", "gatres_small_tough"])
    ap.add_argument("--lr", type=float, default=5e-4)                      # train.py:550
    ap.add_argument("--weight_decay", type=float, default=6e-6)            # train.py:551
    ap.add_argument("--epochs", type=int, default=500)
    ap.add_argument("--mask_rate", type=float, default=0.95)
    ap.add_argument("--batch_size", type=int, default=32)
    ap.add_argument("--patience", type=int, default=100)
    ap.add_argument("--norm_type", default="znorm", choices=["znorm", "minmax", "unused"])
    ap.add_argument("--feature", default="pressure")
    ap.add_argument("--input_path", default=None, help="EPANET .inp of the network")
    ap.add_argument("--dataset_path", default=None, help="zarr v2 store (.zip or directory) written by the reference's generator")
    ap.add_argument("--synthetic", type=int, default=0, help="train on this many smooth synthetic snapshots of a C-Town-shaped network")
    ap.add_argument("--save_path", default="experiments_logs/b200")
    ap.add_argument("--variant", default="b200")
    ap.add_argument("--mask_source", default="device", choices=["device", "numpy"])
    ap.add_argument("--seed", type=int, default=0)
    return ap.parse_args()


def main() -> None:
    args = get_arguments()
    rank, world, local = _dp.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    pg = dist.group.WORLD if world > 1 else None
    args, model = ConfigModels.select_model(args, None)
    if args.synthetic > 0:
        wn = _topology.ctown_shaped()
        data, ei = smooth_synthetic_snapshots(wn, args.synthetic + args.synthetic // 4, seed=args.seed)
        mean, std = float(data[:args.synthetic].mean()), float(data[:args.synthetic].std())
        z = torch.from_numpy(((data - mean) / (std + 1e-8)).astype(np.float32)).to(dev)
        names = list(wn.junctions)
        mk = lambda t: SnapshotSet(t, torch.from_numpy(ei), names, "znorm", mean, std, float(data.min()), float(data.max()))
        train_set, valid_set = mk(z[:args.synthetic]), mk(z[args.synthetic:])
    else:
        if not (args.input_path and args.dataset_path):
            raise SystemExit("give --input_path and --dataset_path, or --synthetic N")
        train_set = SnapshotSet.load(args.input_path, args.dataset_path, args.feature, "train", norm_type=args.norm_type, device=dev)
        valid_set = SnapshotSet.load(args.input_path, args.dataset_path, args.feature, "valid", norm_type=args.norm_type,
                                     mean=train_set.mean, std=train_set.std, min=train_set.min, max=train_set.max, device=dev)
    out = fit(model, train_set, valid_set, args.batch_size, args.epochs, args.lr, args.weight_decay, args.mask_rate,
              args.patience, args.save_path, args.variant, args.mask_source, args.seed, pg)
    if rank == 0:
        b = out["best"]
        print(f"best epoch {b['epoch']}: val_loss {b['loss']:.5f}  " + "  ".join(f"{k} {v:.4f}" for k, v in b["metrics"].items()))


if __name__ == "__main__":
    main()
