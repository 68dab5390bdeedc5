"""Training driver (mirror of the reference's train.py loop on the B200 kernels)."""
import os

import numpy as np
import pytest
import torch

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "caller_ref.npz"))


def test_early_stopping_equals_reference():
    """stop decisions of utils/early_stopping.py (real reference, tests/golden/make_golden_caller.py)"""
    from gnn_pressure_estimation_b200.train import EarlyStopping
    seq = GOLD["early_stopping/seq"].tolist()
    for pat in (0, 1, 3):
        es = EarlyStopping(patience=pat)
        assert [es.step(v) for v in seq] == GOLD[f"early_stopping/patience{pat}"].tolist(), pat


def test_synthetic_snapshots_are_smooth_on_the_graph():
    from gnn_pressure_estimation_b200 import topology
    from gnn_pressure_estimation_b200.train import smooth_synthetic_snapshots
    data, ei = smooth_synthetic_snapshots(topology.ctown_shaped(), 32, seed=1)
    z = (data - data.mean()) / data.std()
    edge_var = float(((z[:, ei[0]] - z[:, ei[1]]) ** 2).mean())
    assert data.shape == (32, 388) and edge_var < 0.5          # neighbours agree far better than independent nodes (2.0)


def test_adam_state_export_follows_model_parameters_order():
    """torch.optim.Adam numbers its state by position in model.parameters() (att_src, att_dst, bias, lin_src.weight
    per GATConv), not by the flat buffer's order: the exported state must attach to the right parameters, and one
    optimizer step after loading it must run (train.py:436 stores optimizer.state_dict())."""
    from types import SimpleNamespace
    import gnn_pressure_estimation_b200.GraphModels as G
    from gnn_pressure_estimation_b200.train import adam_state_dict
    model = G.GATResMeanConv(num_blocks=2, nc=32)
    ordered = model.ordered_parameters()
    P = sum(p.numel() for p in ordered)
    exp_avg = torch.arange(P, dtype=torch.float32)                  # value = offset in the flat buffer
    step = SimpleNamespace(model=model, exp_avg=exp_avg, exp_avg_sq=exp_avg * 2 + 1, step_count=torch.tensor([7]),
                           lr=5e-4, betas=(0.9, 0.999), eps=1e-8, wd=6e-6)
    sd = adam_state_dict(step)
    offsets, off = {}, 0
    for p in ordered:
        offsets[id(p)] = off
        off += p.numel()
    params = list(model.parameters())
    assert len(sd["state"]) == len(params) == len(ordered)
    for i, p in enumerate(params):
        st = sd["state"][i]
        assert st["exp_avg"].shape == p.shape and st["exp_avg_sq"].shape == p.shape
        assert float(st["exp_avg"].reshape(-1)[0]) == offsets[id(p)], i
        assert float(st["step"]) == 7.0
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=6e-6)
    opt.load_state_dict(sd)
    for p in params:
        p.grad = torch.ones_like(p)
    opt.step()                                                        # shapes line up -> no broadcasting error
    assert all(opt.state[p]["exp_avg"].shape == p.shape for p in params)


@pytest.mark.gpu
def test_fit_learns_and_checkpoints(tmp_path):
    """a few epochs on smooth synthetic pressures: validation loss falls below the predict-the-mean level, the last
    (smaller) batch is trained on, checkpoints have the reference's keys and reload into a fresh model"""
    import gnn_pressure_estimation_b200.GraphModels as G
    from gnn_pressure_estimation_b200 import topology, train as TR
    from gnn_pressure_estimation_b200.snapshot_store import SnapshotSet
    dev = torch.device("cuda:0")
    wn = topology.ctown_shaped()
    data, ei = TR.smooth_synthetic_snapshots(wn, 250 + 64, seed=3)
    mean, std = float(data[:250].mean()), float(data[:250].std())
    z = torch.from_numpy(((data - mean) / (std + 1e-8)).astype(np.float32)).to(dev)
    mk = lambda t: SnapshotSet(t, torch.from_numpy(ei), list(wn.junctions), "znorm", mean, std, float(data.min()), float(data.max()))
    train_set, valid_set = mk(z[:250]), mk(z[250:])
    torch.manual_seed(0)
    model = G.GATResMeanConv(name="GATResMeanConv_test", num_blocks=3, nc=32)
    out = TR.fit(model, train_set, valid_set, batch_size=32, epochs=8, lr=3e-3, save_path=str(tmp_path), log_every=1)
    hist = out["history"]
    assert set(out["steps"]) == {32, 250 % 32}
    assert int(out["steps"][32].step_count.item()) == 8 * 8            # 7 full + 1 partial batch per epoch
    assert hist[-1]["val_loss"] < 0.9 * hist[0]["val_loss"] and out["best"]["loss"] < 0.9, [h["val_loss"] for h in hist]
    assert all(k in hist[-1] for k in ("tr_mae", "tr_r2", "val_rmse", "val_mynse", "val_time", "val_throughput"))
    best = tmp_path / "best_GATResMeanConv_test_b200.pth"
    last = tmp_path / "last_GATResMeanConv_test_b200.pth"
    assert best.exists() and last.exists()
    fresh = G.GATResMeanConv(name="x", num_blocks=3, nc=32)
    fresh, cp = TR.load_checkpoint(str(best), fresh)
    for k in ("model_state_dict", "optimizer_state_dict", "epoch", "loss", "val_metric_dict", "mean", "std", "min", "max", "norm_type"):
        assert k in cp, k
    opt = torch.optim.Adam(fresh.parameters(), lr=5e-4, weight_decay=6e-6)
    opt.load_state_dict(cp["optimizer_state_dict"])                     # torch.optim.Adam accepts the exported state
    for p in fresh.parameters():
        assert opt.state[p]["exp_avg"].shape == p.shape
        p.grad = torch.zeros_like(p)
    opt.step()
    assert cp["epoch"] == out["best"]["epoch"] and cp["loss"] == pytest.approx(out["best"]["loss"])
    # NumPy-compatible masks drive the same loop
    out2 = TR.fit(G.GATResMeanConv(num_blocks=2, nc=32), train_set, valid_set, batch_size=64, epochs=1, mask_source="numpy",
                  drop_last=True)
    assert len(out2["history"]) == 1 and np.isfinite(out2["history"][0]["tr_loss"])
