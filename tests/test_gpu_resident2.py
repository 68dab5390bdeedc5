"""Second-generation snapshot-resident kernels (csrc/resident2.cu: exchange tensors in distributed shared memory, rows
in a locality order) against the first generation, the layer-by-layer kernels and the CPU oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from gnn_pressure_estimation_b200 import topology as T
from helpers import assert_close, random_directed_graph, rel_err
from oracle import gatres_oracle as O

pytestmark = pytest.mark.gpu
FWD_TOL, GRAD_TOL = 1e-4, 1e-3


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _graph(kind):
    if kind == "tiny":
        ei, names = T.reference_edge_index(T.tiny_network()); return torch.from_numpy(ei), len(names)
    if kind == "ctown":
        ei, names = T.reference_edge_index(T.ctown_shaped()); return torch.from_numpy(ei), len(names)
    n = 97
    return random_directed_graph(n, 400, seed=5), n


def _setup(kind, B, blocks, dev, seed=3):
    import gnn_pressure_estimation_b200.GraphModels as G
    from gnn_pressure_estimation_b200.graph import Topology
    from gnn_pressure_estimation_b200.train_step import TrainStep
    ei, N = _graph(kind)
    ref = O.make_oracle(blocks, 32, seed=seed)
    model = G.GATResMeanConv(num_blocks=blocks, nc=32)
    model.load_state_dict(ref.state_dict())
    model = model.to(dev)
    topo = Topology.build(ei.to(dev), N)
    ts = TrainStep(model, topo, B, max(1, int(N * 0.95)), use_graph=False)
    x, y, mask = O.synthetic_snapshots(N, B, seed=11)
    ts.load_inputs(y.to(dev), y.to(dev), mask.to(dev))
    return ei, N, ref, model, topo, ts, (x, y, mask)


@pytest.fixture
def knobs():
    from gnn_pressure_estimation_b200 import _lib
    lib = _lib.load()
    prev = (lib.gatres_set_resident_dsm(-1), lib.gatres_set_resident_cluster(-1), lib.gatres_set_resident_max_batch(-1),
            lib.gatres_set_resident_barrier(-1), lib.gatres_set_resident_tc(-1))
    yield lib
    lib.gatres_set_resident_tc(prev[4])
    lib.gatres_set_resident_dsm(prev[0])
    lib.gatres_set_resident_cluster(prev[1])
    lib.gatres_set_resident_max_batch(prev[2])
    lib.gatres_set_resident_barrier(prev[3])


def test_locality_plan_is_a_relabelled_copy_of_the_csr(dev):
    from gnn_pressure_estimation_b200.graph import Topology
    for kind in ("tiny", "ctown", "directed"):
        ei, N = _graph(kind)
        topo = Topology.build(ei.to(dev), N)
        p = topo.plan
        perm = p.perm.cpu().numpy()
        assert sorted(perm.tolist()) == list(range(N))
        for (rp, col), (rpp, colp) in (((topo.rowptr, topo.col), (p.rowptr, p.col)), ((topo.rowptr_t, topo.col_t), (p.rowptr_t, p.col_t))):
            rp, col, rpp, colp = (t.cpu().numpy() for t in (rp, col, rpp, colp))
            assert rpp[-1] == rp[-1]
            for r in range(N):
                i = perm[r]
                assert np.array_equal(perm[colp[rpp[r]:rpp[r + 1]]], col[rp[i]:rp[i + 1]]), (kind, r)   # same entries, same order
        assert p.ecap[0] == topo.E1 and p.ecap[3] <= p.ecap[2] <= p.ecap[1] <= p.ecap[0]
    # the point of the plan: neighbours share a CTA (C-Town-shaped network with random node ids: 11 % before)
    ei, N = _graph("ctown")
    inv = np.empty(N, dtype=np.int64)
    inv[Topology.build(ei.to(dev), N).plan.perm.cpu().numpy()] = np.arange(N)
    R = -(-N // 8)
    assert float((inv[ei[0].numpy()] // R == inv[ei[1].numpy()] // R).mean()) > 0.9


@pytest.mark.parametrize("tc", [1, 0])          # projections on tcgen05 (default where the tiles fit) / mma.sync
@pytest.mark.parametrize("cluster", [0, 1, 2, 4, 8])
@pytest.mark.parametrize("kind,B,blocks", [("tiny", 5, 2), ("directed", 3, 3), ("ctown", 4, 4), ("ctown", 32, 15), ("ctown", 40, 2)])
def test_dsm_inference_forward(kind, B, blocks, cluster, tc, dev, knobs):
    from gnn_pressure_estimation_b200 import _lib
    ei, N, ref, model, topo, ts, (x, y, mask) = _setup(kind, B, blocks, dev)
    d, p, s = C.byref(ts.desc), _lib.ptr, _lib.stream
    lib = knobs
    lib.gatres_set_resident_max_batch(1 << 30)
    lib.gatres_set_resident_cluster(cluster)
    lib.gatres_set_resident_tc(tc)
    scratch = torch.empty(int(lib.gatres_scratch_floats(d, 0)), device=dev)
    _lib.call("gatres_apply_mask", p(ts.x), p(ts.mask), p(ts.xm), ts.M, s())
    outs = {}
    for dsm in (1, 0):
        lib.gatres_set_resident_dsm(dsm)
        n0 = lib.gatres_launch_count()
        ts.out.fill_(float("nan"))
        _lib.call("gatres_forward", d, p(ts.flat), p(ts.xm), p(ts.out), None, p(scratch), s())
        assert lib.gatres_launch_count() - n0 == 1                     # one resident launch either way
        outs[dsm] = ts.out.clone()
    out_ref = ref(x, O.collate_edge_index(ei, N, B)).reshape(-1)
    assert_close(outs[1], out_ref, FWD_TOL, "dsm forward vs oracle")
    assert rel_err(outs[1], outs[0]) < 2e-5, "dsm vs first-generation resident kernel"


@pytest.mark.parametrize("tc", [1, 0])
@pytest.mark.parametrize("barrier", [2, 0, 1])
@pytest.mark.parametrize("cluster", [0, 4, 8])
@pytest.mark.parametrize("kind,B,blocks", [("tiny", 5, 2), ("directed", 3, 3), ("ctown", 4, 4), ("ctown", 32, 15)])
def test_dsm_training_pair_matches_first_generation_and_oracle(kind, B, blocks, cluster, barrier, tc, dev, knobs):
    """forward(training) + backward through the C ABI with the DSMEM kernels, against the first-generation resident
    kernels (same inputs) and against the CPU oracle (forward 1e-4, gradients 1e-3)"""
    from gnn_pressure_estimation_b200 import _lib
    ei, N, ref, model, topo, ts, (x, y, mask) = _setup(kind, B, blocks, dev)
    d, p, s = C.byref(ts.desc), _lib.ptr, _lib.stream
    lib = knobs
    lib.gatres_set_resident_max_batch(1 << 30)
    lib.gatres_set_resident_cluster(cluster)
    lib.gatres_set_resident_barrier(barrier)
    lib.gatres_set_resident_tc(tc)
    res = {}
    for dsm in (1, 0):
        lib.gatres_set_resident_dsm(dsm)
        _lib.call("gatres_apply_mask", p(ts.x), p(ts.mask), p(ts.xm), ts.M, s())
        ts.saved.fill_(float("nan"))
        n0 = lib.gatres_launch_count()
        _lib.call("gatres_forward", d, p(ts.flat), p(ts.xm), p(ts.out), p(ts.saved), p(ts.scratch), s())
        _lib.call("gatres_masked_mse", p(ts.out), p(ts.y), p(ts.mask), ts.M, ts.count, p(ts.d_out), p(ts.loss), p(ts._loss_part), s())
        _lib.call("gatres_backward", d, p(ts.flat), p(ts.xm), p(ts.saved), p(ts.d_out), None, p(ts.grads), p(ts.scratch), s())
        res[dsm] = (ts.out.clone(), ts.grads.clone(), float(ts.loss), lib.gatres_launch_count() - n0)
    eib = O.collate_edge_index(ei, N, B)
    out_ref, loss_ref, grads_ref = O.train_step_loss_and_grads(ref, x, y, mask, eib)
    assert res[1][3] == 4, "forward stack, loss (2 kernels), backward stack"
    assert_close(res[1][0], out_ref.reshape(-1), FWD_TOL, "dsm forward(training) vs oracle")
    assert abs(res[1][2] - float(loss_ref)) <= 2e-4 * abs(float(loss_ref))
    flat_ref = torch.cat([grads_ref[k].reshape(-1) for k in
                          ["lin0.weight", "lin0.bias"] +
                          [f"blocks.{k}.{c}.{t}" for k in range(blocks) for c in ("conv1", "conv2")
                           for t in ("lin_src.weight", "att_src", "att_dst", "bias")] + ["lin1.weight", "lin1.bias"]])
    scale = float(flat_ref.abs().max())
    assert float((res[1][1].cpu() - flat_ref).abs().max()) <= GRAD_TOL * scale, "dsm gradients vs oracle"
    assert float((res[1][1] - res[0][1]).abs().max()) <= 1e-4 * scale, "dsm vs first-generation gradients"
    assert rel_err(res[1][0], res[0][0]) < 2e-5


def test_dsm_backward_in_block_ranges_equals_one_call(dev, knobs):
    """the split backward of the NCCL-overlap path (gatres_backward_range) hands the running gradient from range to
    range through the scratch buffer"""
    from gnn_pressure_estimation_b200 import _lib
    ei, N, ref, model, topo, ts, _ = _setup("ctown", 8, 6, dev)
    d, p, s = C.byref(ts.desc), _lib.ptr, _lib.stream
    lib = knobs
    lib.gatres_set_resident_dsm(1)
    _lib.call("gatres_apply_mask", p(ts.x), p(ts.mask), p(ts.xm), ts.M, s())
    _lib.call("gatres_forward", d, p(ts.flat), p(ts.xm), p(ts.out), p(ts.saved), p(ts.scratch), s())
    _lib.call("gatres_masked_mse", p(ts.out), p(ts.y), p(ts.mask), ts.M, ts.count, p(ts.d_out), p(ts.loss), p(ts._loss_part), s())
    _lib.call("gatres_backward", d, p(ts.flat), p(ts.xm), p(ts.saved), p(ts.d_out), None, p(ts.grads), p(ts.scratch), s())
    whole = ts.grads.clone()
    ts.grads.fill_(float("nan"))
    n0 = lib.gatres_launch_count()
    for hi, lo in ((5, 4), (3, 1), (0, 0)):
        _lib.call("gatres_backward_range", d, p(ts.flat), p(ts.xm), p(ts.saved), p(ts.d_out), None, p(ts.grads), p(ts.scratch), hi, lo, s())
    assert lib.gatres_launch_count() - n0 == 3
    assert float((ts.grads - whole).abs().max()) <= 1e-5 * float(whole.abs().max())

