"""GPU tests of the topology cache (content-validated, never stale), self-loop templates (GATConv drops and re-adds
self loops, SimpleConv(mean) keeps them: SURVEY A.2 step 2 / A.3, call sites GraphModels.py:464-466) and the sibling
models of SURVEY 8f rank 4 (GATConvNet GraphModels.py:15-46, GResBlockConv :548-561) against the CPU oracle."""
import numpy as np
import pytest
import torch

from gnn_pressure_estimation_b200 import topology as T
from helpers import assert_close, random_directed_graph
from oracle import gatres_oracle as O
from oracle import topology_oracle as TO

pytestmark = pytest.mark.gpu
FWD_TOL, GRAD_TOL = 1e-4, 1e-3


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def G():
    import gnn_pressure_estimation_b200.GraphModels as G_
    return G_


def _ctown():
    ei, names = T.reference_edge_index(T.ctown_shaped())
    return torch.from_numpy(ei), len(names)


def _grad_check(model, ref, what):
    gref = {k: q.grad for k, q in ref.named_parameters()}
    floor = 1e-3 * max(float(q.norm()) for q in gref.values())
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        err = float((p.grad.cpu() - gref[k]).abs().max())
        assert err <= GRAD_TOL * max(float(gref[k].abs().max()), floor), f"{what}: {k} {err}"


# ----------------------------------------------------------------------------- topology cache
def test_same_shape_other_wiring_is_recomputed_not_poisoned(dev, G):
    """two collated batches with identical (rows, columns) but different composition — the reference's multi-network
    shuffle case (utils/DataLoader.py:120-129): each gets the right answer, and going back to the first one still
    works (round 1: the second was NaN and so was everything after it)"""
    ei, n = _ctown()
    B = 3
    ref = O.make_oracle(2, 32, seed=3)
    model = G.GATResMeanConv(num_blocks=2, nc=32)
    model.load_state_dict(ref.state_dict())
    model = model.to(dev)
    good = O.collate_edge_index(ei, n, B)
    other = good.clone()
    other[0, 100] = (other[0, 100] + 7) % n                       # same shape, another wiring (no longer replicated)
    x = torch.randn(B * n, 1, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        for ei_b in (good, other, good, other):
            out = model(x.to(dev), ei_b.to(dev))
            assert_close(out, ref(x, ei_b), FWD_TOL, "forward after a topology switch")
    cache = model._topologies
    assert len(cache._by_shape[(B * n, good.size(1), 0)]) == 2     # both compositions cached under one shape


def test_validated_batch_object_costs_no_sync(dev, G):
    ei, n = _ctown()
    model = G.GATResMeanConv(num_blocks=1, nc=32).to(dev)
    eib = O.collate_edge_index(ei, n, 4).to(dev)
    x = torch.randn(4 * n, 1, device=dev)
    with torch.no_grad():
        model(x, eib)
        s0 = model._topologies.syncs
        for _ in range(5):
            model(x, eib)
        assert model._topologies.syncs == s0                      # same tensor, same version: nothing to re-check
        eib[0, 0] = eib[0, 0]                                      # an in-place write bumps the version
        model(x, eib)
        assert model._topologies.syncs == s0 + 1
        fresh = eib.clone()                                        # a new tensor with equal content: one check, a hit
        model(x, fresh)
        assert model._topologies.syncs == s0 + 2
        assert len(model._topologies._by_shape[(4 * n, eib.size(1), 0)]) == 1


def test_unfused_modules_do_not_share_stale_topologies(dev, G):
    """GATConv / SimpleConv modules resolve through a process-wide cache: a second graph of the same shape must not
    run on the first one's CSR"""
    n = 120
    e1 = random_directed_graph(n, 500, seed=1)
    e2 = random_directed_graph(n, e1.size(1) + 40, seed=2)[:, :e1.size(1)]
    assert e1.shape == e2.shape and not torch.equal(e1, e2)
    conv = G.GATConv(32, 32, heads=2, concat=True).to(dev)
    x = torch.randn(n, 32, generator=torch.Generator().manual_seed(0))
    for ei in (e1, e2, e1):
        ref = O.gat_conv(x, ei, conv.lin_src.weight.detach().cpu(), conv.att_src.detach().cpu(),
                         conv.att_dst.detach().cpu(), conv.bias.detach().cpu(), 2, True)
        with torch.no_grad():
            assert_close(conv(x.to(dev), ei.to(dev)), ref, FWD_TOL, "GATConv on a switched graph")


# ----------------------------------------------------------------------------- self loops
def _with_self_loops(ei, n, seed):
    rng = np.random.RandomState(seed)
    loops = torch.from_numpy(rng.choice(n, size=max(2, n // 6), replace=False).astype(np.int64))
    loops = torch.cat([loops, loops[:2]])                          # two nodes carry a DOUBLE self loop
    cols = torch.cat([ei, torch.stack([loops, loops])], dim=1)
    return cols[:, torch.from_numpy(rng.permutation(cols.size(1)))]   # interleaved with the ordinary edges


def test_csr_views_of_a_template_with_self_loops(dev):
    from gnn_pressure_estimation_b200.graph import Topology
    n = 97
    ei = _with_self_loops(random_directed_graph(n, 300, seed=5), n, seed=6)
    topo = Topology.build(ei.to(dev), n)
    k = int((ei[0] == ei[1]).sum())
    assert topo.dropped_self_loops == k and not topo.shares_one_csr and topo.E1 == ei.size(1) - k + n
    keep = ei[:, ei[0] != ei[1]].numpy()
    rp, col = TO.csr_by_target(keep, n)                            # GATConv's view: loops dropped, one appended
    assert np.array_equal(topo.rowptr.cpu().numpy(), rp) and np.array_equal(topo.col.cpu().numpy(), col)
    rpm, colm, rptm, coltm = (t.cpu().numpy() for t in topo.mean_view())
    # SimpleConv's view: stable sort of the ORIGINAL list by target, plus the trailing entry the kernels skip
    order = np.argsort(ei[1].numpy(), kind="stable")
    for i in range(n):
        want = ei[0].numpy()[order][ei[1].numpy()[order] == i]
        assert np.array_equal(colm[rpm[i]:rpm[i + 1] - 1], want), i
    order_t = np.argsort(ei[0].numpy(), kind="stable")
    for j in range(n):
        want = ei[1].numpy()[order_t][ei[0].numpy()[order_t] == j]
        assert np.array_equal(coltm[rptm[j]:rptm[j + 1] - 1], want), j


@pytest.mark.parametrize("graph,B", [("ctown", 3), ("directed", 2)])
def test_model_on_template_with_self_loops_matches_oracle(graph, B, dev, G):
    if graph == "ctown":
        ei, n = _ctown()
    else:
        n = 97
        ei = random_directed_graph(n, 400, seed=15)
    ei = _with_self_loops(ei, n, seed=7)
    ref = O.make_oracle(3, 32, seed=5)
    model = G.GATResMeanConv(num_blocks=3, nc=32)
    model.load_state_dict(ref.state_dict())
    model = model.to(dev)
    x, y, mask = O.synthetic_snapshots(n, B, seed=3)
    eib = O.collate_edge_index(ei, n, B)
    out_ref, loss_ref, grads_ref = O.train_step_loss_and_grads(ref, x, y, mask, eib)
    out = model(x.to(dev), eib.to(dev), None, None)
    md = mask.to(dev)
    torch.nn.functional.mse_loss(out[md], y.to(dev)[md]).backward()
    assert_close(out, out_ref, FWD_TOL, "self-loop template forward")
    floor = 1e-3 * max(float(g.norm()) for g in grads_ref.values())
    for k, p in model.named_parameters():
        g = grads_ref[k]
        assert float((p.grad.cpu() - g).abs().max()) <= GRAD_TOL * max(float(g.abs().max()), floor), k
    # the self loops matter: dropping them from the mean changes the answer well beyond the tolerance
    no_loops = O.collate_edge_index(ei[:, ei[0] != ei[1]], n, B)
    assert float((ref(x, no_loops) - out_ref).abs().max()) > 100 * FWD_TOL * float(out_ref.abs().max())


def test_train_step_rejects_self_loop_templates(dev, G):
    from gnn_pressure_estimation_b200.graph import Topology
    from gnn_pressure_estimation_b200.train_step import TrainStep
    n = 50
    ei = _with_self_loops(random_directed_graph(n, 150, seed=1), n, seed=2)
    topo = Topology.build(ei.to(dev), n)
    with pytest.raises(NotImplementedError):
        TrainStep(G.GATResMeanConv(num_blocks=1, nc=32).to(dev), topo, 2, 10)


# ----------------------------------------------------------------------------- sibling models
@pytest.mark.parametrize("train", [False, True])
def test_gatconvnet_matches_oracle(train, dev, G):
    ei, n = _ctown()
    B = 2
    net = dict(input_dim=1, hidden_dim=32, heads=2, out_dim=1, num_layers=4)
    torch.manual_seed(11)
    ref = O.GATConvNetOracle(net)
    model = G.GATConvNet(net)
    model.load_state_dict(ref.state_dict())
    model = model.to(dev)
    eib = O.collate_edge_index(ei, n, B)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B * n, 1, generator=g)
    y = torch.rand(B * n, 1, generator=g)
    masks = [(torch.rand(B * n, 64, generator=g) > 0.5).float() for _ in range(3)] if train else None
    ref.train(train), model.train(train)
    out_ref = ref(x, eib, None, masks)
    torch.nn.functional.mse_loss(out_ref, y).backward()
    out = model(x.to(dev), eib.to(dev), None, [m.to(dev) for m in masks] if train else None)
    torch.nn.functional.mse_loss(out, y.to(dev)).backward()
    assert_close(out, out_ref, FWD_TOL, "GATConvNet forward")
    _grad_check(model, ref, "GATConvNet")


def test_gresblockconv_matches_oracle(dev, G):
    ei, n = _ctown()
    B = 3
    torch.manual_seed(4)
    ref = O.OracleGResBlockConv(32, 32, 32)
    with torch.no_grad():
        ref.conv1.bias.uniform_(-0.1, 0.1), ref.conv2.bias.uniform_(-0.1, 0.1)
    blk = G.GResBlockConv(32, 32, 32)
    blk.load_state_dict(ref.state_dict())
    blk = blk.to(dev)
    eib = O.collate_edge_index(ei, n, B)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B * n, 32, generator=g)
    go = torch.randn(B * n, 32, generator=g)
    xr = x.clone().requires_grad_()
    ref(xr, eib, None).backward(go)
    xd = x.to(dev).requires_grad_()
    out = blk(xd, eib.to(dev), None)
    out.backward(go.to(dev))
    assert_close(out, ref(x, eib, None), FWD_TOL, "GResBlockConv forward")
    assert_close(xd.grad, xr.grad, GRAD_TOL, "GResBlockConv dx")
    _grad_check(blk, ref, "GResBlockConv")


def test_hand_composed_gatres_uses_library_linears(dev, G):
    """lin0 / lin1 through Linear.forward run the encoder / decoder kernels (no cuBLAS on a hand-composed GATRes)"""
    from gnn_pressure_estimation_b200 import _lib
    lib = _lib.load()
    ref = O.make_oracle(1, 32, seed=1)
    model = G.GATResMeanConv(num_blocks=1, nc=32)
    model.load_state_dict(ref.state_dict())
    model = model.to(dev)
    x = torch.randn(64, 1, generator=torch.Generator().manual_seed(2))
    n0 = lib.gatres_launch_count()
    h = model.lin0(x.to(dev))
    out = model.lin1(h)
    assert lib.gatres_launch_count() - n0 == 2
    out.sum().backward()
    hr = ref.lin0(x)
    outr = ref.lin1(hr)
    outr.sum().backward()
    assert_close(out, outr, FWD_TOL, "lin1(lin0(x))")
    for name in ("lin0.weight", "lin0.bias", "lin1.weight", "lin1.bias"):
        a = dict(model.named_parameters())[name].grad
        b = dict(ref.named_parameters())[name].grad
        assert_close(a, b, GRAD_TOL, name)
