"""GPU parity at the exact shapes of the BASELINE.json configurations, against the CPU oracle:

  configs[1]  gatres_small, 15 blocks x B = 32 on C-Town through TrainStep (the bench step: resident cluster kernels)
  configs[3]  gatres_large (25 blocks x 128 channels) at B = 64 (layer path: wide tcgen05 projections, tile / gather
              aggregation as the dispatcher picks them at this batch)
  configs[4]  gatres_small on the scaled synthetic WDN (100 000 junctions): gather aggregation with the CSR in L2

Tolerances as everywhere (north_star): forward <= 1e-4, gradients <= 1e-3 relative, fp32; gradients of a tensor are
held to max(own scale, 1e-3 x the largest parameter-gradient norm) (the att_dst cancellation floor).
"""
import pytest
import torch

from gnn_pressure_estimation_b200 import topology as T
from helpers import assert_close
from oracle import gatres_oracle as O

pytestmark = pytest.mark.gpu
FWD_TOL, GRAD_TOL = 1e-4, 1e-3


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def G():
    import gnn_pressure_estimation_b200.GraphModels as G_
    return G_


def _pair(G, dev, nb, nc, seed=0):
    ref = O.make_oracle(nb, nc, seed=seed)
    m = G.GATResMeanConv(num_blocks=nb, nc=nc)
    m.load_state_dict(ref.state_dict())
    return m.to(dev), ref


def _check_grads(named_grads, grads_ref, what):
    floor = 1e-3 * max(float(g.norm()) for g in grads_ref.values())
    worst = 0.0
    for k, g in named_grads:
        r = grads_ref[k]
        assert g.shape == r.shape, k
        e = float((g.cpu() - r).abs().max()) / max(float(r.abs().max()), floor)
        worst = max(worst, e)
        assert e <= GRAD_TOL, f"{what}: {k} rel err {e:.3e}"
    return worst


def test_headline_step_15_blocks_batch_32_matches_oracle(dev, G):
    """the exact BENCH shape: 15 blocks, 32 snapshots of the C-Town-shaped graph, one captured TrainStep replay"""
    from gnn_pressure_estimation_b200.train_step import TrainStep
    ei_np, names = T.reference_edge_index(T.ctown_shaped())
    ei, N, B = torch.from_numpy(ei_np), len(names), 32
    model, ref = _pair(G, dev, 15, 32)
    topo = model.set_topology(ei.to(dev), N)
    ts = TrainStep(model, topo, B, mask_count_per_snapshot=int(N * 0.95), use_graph=True)
    ts.capture()
    assert ts.kernels_per_step < 20, "B = 32 must take the snapshot-resident path (7 launches per step)"
    x, y, mask = O.synthetic_snapshots(N, B, seed=77)
    eib = O.collate_edge_index(ei, N, B)
    out_ref, loss_ref, grads_ref = O.train_step_loss_and_grads(ref, x, y, mask, eib)
    loss = ts.step(y.to(dev), y.to(dev), mask.to(dev))
    torch.cuda.synchronize()
    assert_close(ts.out.view(-1, 1), out_ref, FWD_TOL, "forward, 15 blocks x B=32")
    assert abs(float(loss) - float(loss_ref)) <= 2e-4 * abs(float(loss_ref))
    named, off = [], 0
    names_in_flat_order = ["lin0.weight", "lin0.bias"]
    for k in range(15):
        for c in ("conv1", "conv2"):
            names_in_flat_order += [f"blocks.{k}.{c}.lin_src.weight", f"blocks.{k}.{c}.att_src", f"blocks.{k}.{c}.att_dst",
                                    f"blocks.{k}.{c}.bias"]
    names_in_flat_order += ["lin1.weight", "lin1.bias"]
    for name, p in zip(names_in_flat_order, model.ordered_parameters()):
        named.append((name, ts.grads[off:off + p.numel()].view(p.shape)))
        off += p.numel()
    assert off == ts.grads.numel()
    _check_grads(named, grads_ref, "TrainStep gradients")


def test_scaled_wdn_model_matches_oracle(dev, G):
    """configs[4]: gatres_small (15 blocks) on the 100 000-junction synthetic network, training and inference"""
    ei_np, names = T.reference_edge_index(T.scaled_wdn())
    ei, N, B = torch.from_numpy(ei_np), len(names), 2
    assert N == 100_000
    model, ref = _pair(G, dev, 15, 32)
    x, y, mask = O.synthetic_snapshots(N, B, seed=5)
    eib = O.collate_edge_index(ei, N, B)
    out_ref, loss_ref, grads_ref = O.train_step_loss_and_grads(ref, x, y, mask, eib)
    eibd, md = eib.to(dev), mask.to(dev)
    out = model(x.to(dev), eibd, None, None)
    loss = torch.nn.functional.mse_loss(out[md], y.to(dev)[md])
    loss.backward()
    assert_close(out, out_ref, FWD_TOL, "scaled WDN forward")
    assert abs(float(loss) - float(loss_ref)) <= 2e-4 * abs(float(loss_ref))
    _check_grads([(k, p.grad) for k, p in model.named_parameters()], grads_ref, "scaled WDN gradients")
    with torch.no_grad():
        assert_close(model(x.to(dev), eibd), out_ref, FWD_TOL, "scaled WDN inference forward")


def test_scaled_wdn_train_step_matches_oracle(dev, G):
    """the same network through the captured TrainStep (bench leg configs[4] train): loss and gradients"""
    from gnn_pressure_estimation_b200.train_step import TrainStep
    ei_np, names = T.reference_edge_index(T.scaled_wdn())
    ei, N, B = torch.from_numpy(ei_np), len(names), 2
    model, ref = _pair(G, dev, 3, 32, seed=2)
    topo = model.set_topology(ei.to(dev), N)
    ts = TrainStep(model, topo, B, mask_count_per_snapshot=int(N * 0.95), use_graph=True)
    ts.capture()
    x, y, mask = O.synthetic_snapshots(N, B, seed=6)
    out_ref, loss_ref, grads_ref = O.train_step_loss_and_grads(ref, x, y, mask, O.collate_edge_index(ei, N, B))
    loss = ts.step(y.to(dev), y.to(dev), mask.to(dev))
    assert_close(ts.out.view(-1, 1), out_ref, FWD_TOL, "scaled WDN TrainStep forward")
    assert abs(float(loss) - float(loss_ref)) <= 2e-4 * abs(float(loss_ref))
    flat_ref = torch.cat([grads_ref[k].reshape(-1) for k in
                          ["lin0.weight", "lin0.bias"] +
                          [f"blocks.{k}.{c}.{t}" for k in range(3) for c in ("conv1", "conv2")
                           for t in ("lin_src.weight", "att_src", "att_dst", "bias")] + ["lin1.weight", "lin1.bias"]])
    err = float((ts.grads.cpu() - flat_ref).abs().max()) / float(flat_ref.abs().max())
    assert err <= GRAD_TOL, err


@pytest.mark.parametrize("B", [64, 80])
def test_large_model_layer_path_matches_oracle(B, dev, G):
    """configs[3] model: gatres_large = 25 blocks x 128 channels, at batches that take the large-batch kernels
    (>= 64 snapshots: >= 24 832 rows per launch; 80 is not a multiple of the 128-row GEMM tiles per snapshot)"""
    ei_np, names = T.reference_edge_index(T.ctown_shaped())
    ei, N = torch.from_numpy(ei_np), len(names)
    model, ref = _pair(G, dev, 25, 128)
    x, y, mask = O.synthetic_snapshots(N, B, seed=9)
    eib = O.collate_edge_index(ei, N, B)
    out_ref, loss_ref, grads_ref = O.train_step_loss_and_grads(ref, x, y, mask, eib)
    md = mask.to(dev)
    out = model(x.to(dev), eib.to(dev), None, None)
    loss = torch.nn.functional.mse_loss(out[md], y.to(dev)[md])
    loss.backward()
    assert_close(out, out_ref, FWD_TOL, "gatres_large forward")
    _check_grads([(k, p.grad) for k, p in model.named_parameters()], grads_ref, "gatres_large gradients")
    with torch.no_grad():
        assert_close(model(x.to(dev), eib.to(dev)), out_ref, FWD_TOL, "gatres_large inference forward")


@pytest.mark.parametrize("path", ["layer", "resident-waves"])
def test_small_model_large_batch_tile_path_matches_oracle(path, dev, G):
    """configs[2] regime at B = 96 (37 248 rows).  "layer": the TMA snapshot-tile aggregation kernels, the tcgen05
    projections and the fused projection backward (what batches beyond ~200 snapshots take; forced here).
    "resident-waves": the default at this size — the snapshot-resident cluster kernels with 8 CTAs per snapshot, 96
    clusters on 37 cluster slots, i.e. three waves of one launch."""
    from gnn_pressure_estimation_b200 import _lib
    lib = _lib.load()
    ei_np, names = T.reference_edge_index(T.ctown_shaped())
    ei, N, B = torch.from_numpy(ei_np), len(names), 96
    model, ref = _pair(G, dev, 15, 32)
    x, y, mask = O.synthetic_snapshots(N, B, seed=10)
    eib = O.collate_edge_index(ei, N, B)
    out_ref, loss_ref, grads_ref = O.train_step_loss_and_grads(ref, x, y, mask, eib)
    md = mask.to(dev)
    prev = lib.gatres_set_resident_max_batch(0) if path == "layer" else None
    try:
        n0 = lib.gatres_launch_count()
        out = model(x.to(dev), eib.to(dev), None, None)
        torch.nn.functional.mse_loss(out[md], y.to(dev)[md]).backward()
        launches = lib.gatres_launch_count() - n0
    finally:
        if prev is not None:
            lib.gatres_set_resident_max_batch(prev)
    assert (launches > 100) if path == "layer" else (launches < 10), launches
    assert_close(out, out_ref, FWD_TOL, "gatres_small B=96 forward")
    _check_grads([(k, p.grad) for k, p in model.named_parameters()], grads_ref, "gatres_small B=96 gradients")
