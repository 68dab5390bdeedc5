"""GPU parity tests: the CUDA path (through torch.ops.gatres -> C ABI) against the
CPU oracle on identical inputs and weights, and against the committed goldens.

Tolerances (BASELINE.json north_star): indexing/CSR bit-exact; forward <= 1e-4
relative; gradients <= 1e-3 relative, fp32.  rel = max|a-b| / max|b| per tensor.
"""
import numpy as np
import pytest
import torch

from gnn_pressure_estimation_b200 import topology as T
from helpers import assert_close, load_case, random_directed_graph, rel_err
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


def _topology(ei_cpu, n, dev):
    from gnn_pressure_estimation_b200.graph import Topology
    return Topology.build(ei_cpu.to(dev), n)


# ------------------------------------------------------------------ indexing
@pytest.mark.parametrize("case", ["tiny", "ctown", "scaled", "directed"])
def test_csr_bit_exact(case, dev):
    if case == "tiny":
        ei, names = T.reference_edge_index(T.tiny_network()); n = len(names)
    elif case == "ctown":
        ei, names = T.reference_edge_index(T.ctown_shaped()); n = len(names)
    elif case == "scaled":
        ei, names = T.reference_edge_index(T.scaled_wdn(20_000, 23_000, seed=1)); n = len(names)
    else:
        n = 301
        ei = random_directed_graph(n, 2000, seed=11).numpy()
    topo = _topology(torch.from_numpy(ei), n, dev)
    rp, col = TO.csr_by_target(ei, n)
    rpt, colt = TO.csr_by_source(ei, n)
    assert topo.E1 == ei.shape[1] + n
    assert np.array_equal(topo.rowptr.cpu().numpy(), rp) and np.array_equal(topo.col.cpu().numpy(), col)
    assert np.array_equal(topo.rowptr_t.cpu().numpy(), rpt) and np.array_equal(topo.col_t.cpu().numpy(), colt)


def test_csr_matches_golden_fixture(dev, golden_dir):
    g = np.load(f"{golden_dir}/topo_ctown_shaped.npz")
    topo = _topology(torch.from_numpy(g["edge_index"]), int(g["n"]), dev)
    for got, key in ((topo.rowptr, "rowptr"), (topo.col, "col"), (topo.rowptr_t, "rowptr_t"), (topo.col_t, "col_t")):
        assert np.array_equal(got.cpu().numpy(), g[key]), key


def test_csr_rejects_bad_ids_and_splits_self_loops(dev):
    from gnn_pressure_estimation_b200._lib import GatresError
    with pytest.raises(GatresError):
        _topology(torch.tensor([[0, 5], [1, 0]]), 3, dev)
    # a template with a self loop gets two views (GATConv drops it, SimpleConv keeps it): tests/test_gpu_robustness.py
    topo = _topology(torch.tensor([[0, 1, 1], [1, 0, 1]]), 3, dev)
    assert topo.dropped_self_loops == 1 and not topo.shares_one_csr
    assert topo.col.tolist() == [1, 0, 0, 1, 2] and topo.mean_view()[1].tolist() == [1, 0, 0, 1, 1, 2]


def test_replication_check_and_poison(dev, G):
    """eager calls validate the batch content and re-infer on a mismatch (correct answer, general mode); only under
    CUDA-graph capture, where no host read is possible, a mismatching batch poisons the output with NaN"""
    ei, names = T.reference_edge_index(T.ctown_shaped()); n = len(names)
    ei = torch.from_numpy(ei)
    ref = O.make_oracle(1, 32, seed=2)
    model = G.GATResMeanConv(num_blocks=1, nc=32)
    model.load_state_dict(ref.state_dict())
    model = model.to(dev)
    x = torch.randn(4 * n, 1, generator=torch.Generator().manual_seed(3))
    good = O.collate_edge_index(ei, n, 4)
    bad = good.clone()
    bad[0, 2000] = (bad[0, 2000] + 1) % n + 2 * n              # same shape, different wiring
    xd, goodd, badd = x.to(dev), good.to(dev), bad.to(dev)
    with torch.no_grad():
        assert_close(model(xd, goodd), ref(x, good), FWD_TOL, "replicated batch")
        assert_close(model(xd, badd), ref(x, bad), FWD_TOL, "same shape, other wiring: re-inferred, not stale")
        assert_close(model(xd, goodd), ref(x, good), FWD_TOL, "and back")
        # captured: the check is enqueued, its flag is zeroed per replay, NaN on mismatch
        fresh_model = G.GATResMeanConv(num_blocks=1, nc=32).to(dev)
        fresh_model(xd, goodd)
        static_ei = goodd.clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fresh_model(xd, static_ei.clone())
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        capt_ei = static_ei.clone()                             # a tensor the cache has not validated
        with torch.cuda.graph(graph):
            out = fresh_model(xd, capt_ei)
        graph.replay()
        assert torch.isfinite(out).all()
        capt_ei.copy_(badd)
        graph.replay()
        assert torch.isnan(out).all()
        capt_ei.copy_(goodd)
        graph.replay()
        assert torch.isfinite(out).all()                        # the flag does not stick


# ----------------------------------------------------------------- operators
def _layer_inputs(M, fin, H, C, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, fin, generator=g)
    W = torch.randn(H * C, fin, generator=g) * (1.0 / fin ** 0.5)
    a_s = torch.randn(1, H, C, generator=g) * 0.3
    a_d = torch.randn(1, H, C, generator=g) * 0.3
    return x, W, a_s, a_d


GRAPHS = {
    "tiny": lambda: T.reference_edge_index(T.tiny_network()),
    "ctown": lambda: T.reference_edge_index(T.ctown_shaped()),
}


@pytest.fixture(params=["gather-ffma", "tile-tc", "resident"])
def kernel_variant(request):
    """run a test with (a) the gather aggregation kernels + fp32 FFMA projections, (b) the TMA-staged
    snapshot-tile aggregation kernels (default only for batches >= 64) + tcgen05 3xTF32 projections, and
    (c) the snapshot-resident cluster kernels (whole-model calls of small batches; op-level calls are
    unaffected).  (a) and (b) switch the resident path off so that the layer kernels stay covered."""
    from gnn_pressure_estimation_b200 import _lib
    lib = _lib.load()
    fancy = request.param == "tile-tc"
    prev_tile = lib.gatres_set_tile_min_batch(1 if fancy else 1 << 40)
    prev_tc = lib.gatres_set_tensor_core(2 if fancy else 0)
    prev_res = lib.gatres_set_resident_max_batch(256 if request.param == "resident" else 0)
    yield request.param
    lib.gatres_set_tile_min_batch(prev_tile)
    lib.gatres_set_tensor_core(prev_tc)
    lib.gatres_set_resident_max_batch(prev_res)


@pytest.mark.parametrize("graph,B", [("tiny", 1), ("tiny", 5), ("ctown", 3), ("directed", 2)])
@pytest.mark.parametrize("H,C,fin,concat,relu", [(2, 32, 32, True, True), (1, 32, 64, False, False),
                                                 (2, 64, 64, True, True), (1, 64, 128, False, False),
                                                 (2, 128, 128, True, False), (1, 128, 256, False, True)])
def test_gat_conv_forward_backward(graph, B, H, C, fin, concat, relu, dev, kernel_variant):
    from gnn_pressure_estimation_b200 import ops as gops
    if graph == "directed":
        n = 97
        ei = random_directed_graph(n, 400, seed=5)
    else:
        ei_np, names = GRAPHS[graph](); n = len(names); ei = torch.from_numpy(ei_np)
    M = B * n
    x, W, a_s, a_d = _layer_inputs(M, fin, H, C, seed=H * 1000 + C + B)
    bias = torch.randn(H * C if concat else C, generator=torch.Generator().manual_seed(3)) * 0.1
    eib = O.collate_edge_index(ei, n, B)
    leaves = [t.clone().requires_grad_() for t in (x, W, a_s, a_d, bias)]
    ref = O.gat_conv(leaves[0], eib, leaves[1], leaves[2], leaves[3], leaves[4], H, concat)
    if relu:
        ref = ref.relu()
    go = torch.randn(ref.shape, generator=torch.Generator().manual_seed(9))
    ref.backward(go)

    topo = _topology(ei, n, dev)
    cl = [t.detach().clone().to(dev).requires_grad_() for t in (x, W, a_s, a_d, bias)]
    out = gops.gat_conv(cl[0], cl[1], cl[2], cl[3], cl[4], topo, B, H, concat, relu=relu)
    out.backward(go.to(dev))
    assert_close(out, ref, FWD_TOL, "gat_conv forward")
    for name, a, b in zip(("dx", "dW", "datt_src", "datt_dst", "dbias"), cl, leaves):
        assert_close(a.grad, b.grad, GRAD_TOL, name)


@pytest.mark.parametrize("H,C,fin", [(2, 128, 128), (1, 128, 256), (2, 64, 64), (1, 64, 128), (2, 32, 32), (1, 32, 64)])
def test_tensor_core_projection_many_tiles(H, C, fin, dev):
    """tcgen05 3xTF32 projections (whole-K kernel for nc = 32, K-chunked pipeline for the wide shapes) against the
    fp32 FFMA kernel and an fp64 reference, with several 128-row tiles per CTA and a ragged last tile."""
    from gnn_pressure_estimation_b200 import _lib, ops as gops  # noqa: F401
    lib = _lib.load()
    M = 128 * 148 * 5 + 128 * 37 + 5                         # several tiles per persistent CTA (staging-buffer reuse), ragged end
    x, W, a_s, a_d = _layer_inputs(M, fin, H, C, seed=17)
    res = []
    prev = lib.gatres_set_tensor_core(-1)
    try:
        for mode in (0, 2):
            lib.gatres_set_tensor_core(mode)
            h, ss, sd = torch.ops.gatres.linear_att_fwd(x.to(dev), W.to(dev), a_s.reshape(-1).to(dev), a_d.reshape(-1).to(dev), H, C)
            res.append((h.cpu(), ss.cpu(), sd.cpu()))
    finally:
        lib.gatres_set_tensor_core(prev)
    h64 = x.double() @ W.double().T
    s64 = (h64.view(M, H, C) * a_s.double()).sum(-1)
    d64 = (h64.view(M, H, C) * a_d.double()).sum(-1)
    for name, (h, ss, sd) in zip(("ffma", "tcgen05"), res):
        assert_close(h, h64.float(), 5e-6, f"{name} h")          # fp32 accumulation over K <= 256
        assert_close(ss.view(M, H), s64.float(), 1e-5, f"{name} s_src")
        assert_close(sd.view(M, H), d64.float(), 1e-5, f"{name} s_dst")

@pytest.mark.parametrize("H,C,fin,concat", [(2, 32, 32, True), (1, 32, 64, False), (1, 64, 128, False),
                                            # rows wider than a slab: channel-sliced work units (2-D TMA tensor loads)
                                            (2, 64, 64, True), (2, 128, 128, True), (1, 128, 256, False)])
def test_tile_kernels_many_snapshots_per_cta(H, C, fin, concat, dev):
    """333 snapshots = more than two per persistent CTA (148 SMs): the software pipelines of the snapshot-tile kernels
    (double-buffered stages, the two-stage backward that loads one half of a snapshot under the other pass, the
    SimpleConv(mean) backward tile kernel) run several iterations per CTA.  Checked against the gather kernels on the same
    inputs (those are held to the oracle at small sizes above) and, for the first snapshots, against the oracle."""
    from gnn_pressure_estimation_b200 import _lib, ops as gops
    lib = _lib.load()
    ei_np, names = GRAPHS["ctown"](); n = len(names); ei = torch.from_numpy(ei_np)
    B = 333
    M = B * n
    x, W, a_s, a_d = _layer_inputs(M, fin, H, C, seed=77)
    bias = torch.randn(H * C if concat else C, generator=torch.Generator().manual_seed(3)) * 0.1
    go = torch.randn(M, H * C if concat else C, generator=torch.Generator().manual_seed(9))
    topo = _topology(ei, n, dev)
    res = {}
    prev_tile, prev_tc = lib.gatres_set_tile_min_batch(-1), lib.gatres_set_tensor_core(-1)
    try:
        for variant, min_batch in (("tile", 1), ("gather", 1 << 40)):
            lib.gatres_set_tile_min_batch(min_batch)
            lib.gatres_set_tensor_core(0)
            cl = [t.detach().clone().to(dev).requires_grad_() for t in (x, W, a_s, a_d, bias)]
            out = gops.gat_conv(cl[0], cl[1], cl[2], cl[3], cl[4], topo, B, H, concat, relu=True)
            out.backward(go.to(dev))
            res[variant] = [out.detach()] + [t.grad for t in cl]
            if C == 32 and not concat:                      # SimpleConv(mean) backward, model form (mask already applied)
                gm = go.to(dev).contiguous()
                dz = torch.full_like(gm, float("nan"))
                _lib.call("gatres_mean_res_bwd_e1", _lib.ptr(topo.rowptr), _lib.ptr(topo.rowptr_t), _lib.ptr(topo.col_t),
                          topo.E1, _lib.ptr(gm), _lib.ptr(dz), B, n, C, _lib.stream())
                res[variant].append(dz)
    finally:
        lib.gatres_set_tile_min_batch(prev_tile)
        lib.gatres_set_tensor_core(prev_tc)
    for name, a, b in zip(("out", "dx", "dW", "datt_src", "datt_dst", "dbias", "mean_bwd dz"), res["tile"], res["gather"]):
        assert_close(a, b, 2e-5, f"tile vs gather: {name}")
    # first two snapshots against the oracle (forward)
    k = 2 * n
    ref = O.gat_conv(x[:k], O.collate_edge_index(ei, n, 2), W, a_s, a_d, bias, H, concat).relu()
    assert_close(res["tile"][0][:k], ref, FWD_TOL, "tile forward vs oracle")



@pytest.mark.parametrize("switch", ["GATRES_TC_PAIR=1", "GATRES_TC_WIDE2_TMA_STORE=0", "GATRES_TC_WIDE2=0"])
def test_projection_kernel_variants_in_a_subprocess(switch, dev):
    """Kernel-selection switches of the projections are read once per process: the cta_group::2 form (opt-in), the
    register-store epilogue (the fallback when no tensor map can be encoded) and the first-generation kernels run the probe
    of tools/wide_probe.py in a subprocess; its fp64 comparison is held to the tolerance of the in-process test."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    key, val = switch.split("=")
    env = dict(os.environ, **{key: val})
    rows = 128 * 148 * 4 + 128 * 5 + 77                      # odd number of tiles: the last pair is half empty and ragged
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "wide_probe.py"), "--rows", str(rows), "--iters", "2",
                          "--narrow"], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    recs = [json.loads(line) for line in out.stdout.splitlines() if line.startswith("{")]
    assert len(recs) == 6
    for r in recs:
        assert r["err_h"] < 5e-6 and r["err_s_src"] < 1e-5 and r["err_s_dst"] < 1e-5, r


@pytest.mark.parametrize("graph,B", [("tiny", 4), ("ctown", 2), ("directed", 3)])
@pytest.mark.parametrize("C", [32, 64, 128])
def test_mean_res_forward_backward(graph, B, C, dev):
    from gnn_pressure_estimation_b200 import ops as gops
    if graph == "directed":
        n = 97
        ei = random_directed_graph(n, 400, seed=6)
    else:
        ei_np, names = GRAPHS[graph](); n = len(names); ei = torch.from_numpy(ei_np)
    g = torch.Generator().manual_seed(C + B)
    z = torch.randn(B * n, C, generator=g)
    x0 = torch.randn(B * n, C, generator=g)
    eib = O.collate_edge_index(ei, n, B)
    zr, xr = z.clone().requires_grad_(), x0.clone().requires_grad_()
    ref = (O.simple_conv_mean(zr, eib) + xr).relu()
    go = torch.randn(ref.shape, generator=g)
    ref.backward(go)
    topo = _topology(ei, n, dev)
    zc, xc = z.to(dev).requires_grad_(), x0.to(dev).requires_grad_()
    out = gops.mean_res(zc, xc, topo, B)
    out.backward(go.to(dev))
    assert_close(out, ref, 1e-6, "mean_res forward")
    assert_close(zc.grad, zr.grad, 1e-6, "dz")
    assert_close(xc.grad, xr.grad, 1e-6, "dres")
    if graph == "tiny":                                   # isolated node J7: mean part is 0
        assert torch.equal(out.view(B, n, C)[:, 6].cpu(), x0.view(B, n, C)[:, 6].relu())


def test_block_module_matches_oracle(dev, G):
    ei_np, names = T.reference_edge_index(T.ctown_shaped()); n = len(names); ei = torch.from_numpy(ei_np)
    B = 2
    ref = O.make_oracle(1, 32, seed=4).blocks[0]
    blk = G.GResBlockMeanConv(32, 32, 32)
    blk.load_state_dict(ref.state_dict())
    blk = blk.to(dev)
    x = torch.randn(B * n, 32, generator=torch.Generator().manual_seed(1))
    eib = O.collate_edge_index(ei, n, B)
    xr = x.clone().requires_grad_()
    o_ref = ref(xr, eib)
    o_ref.square().sum().backward()
    xc = x.to(dev).requires_grad_()
    o = blk(xc, eib.to(dev))
    o.square().sum().backward()
    assert_close(o, o_ref, FWD_TOL, "block forward")
    assert_close(xc.grad, xr.grad, GRAD_TOL, "block dx")
    for (k, p), q in zip(blk.named_parameters(), ref.parameters()):
        assert_close(p.grad, q.grad, GRAD_TOL, k)


# --------------------------------------------------------------- whole model
def _cuda_model_from_case(c, G, dev):
    ref = O.make_oracle(c["num_blocks"], c["nc"], seed=c["seed"])
    m = G.GATResMeanConv(num_blocks=c["num_blocks"], nc=c["nc"])
    m.load_state_dict(ref.state_dict())
    return m.to(dev), ref


@pytest.mark.parametrize("deterministic", [False, True])
@pytest.mark.parametrize("name", ["tiny_2b_32c_B3", "ctown_small_15b_32c_B8", "ctown_mid_3b_64c_B2",
                                  "ctown_large_25b_128c_B2"])
def test_model_matches_golden(name, deterministic, dev, G, kernel_variant):
    """reference caller semantics (train.py:174-185): mask applied by the caller, MSE on masked nodes."""
    c = load_case(name)
    model, _ = _cuda_model_from_case(c, G, dev)
    model.deterministic = deterministic
    eib = O.collate_edge_index(c["edge_index"], c["N"], c["B"]).to(dev)
    mask = c["mask"].to(dev)
    out = model(c["x"].to(dev), eib, None, None)
    loss = torch.nn.functional.mse_loss(out[mask], c["y"].to(dev)[mask])
    loss.backward()
    assert out.shape == c["out"].shape
    assert_close(out, c["out"], FWD_TOL, "forward")
    assert abs(float(loss) - float(c["loss"])) <= FWD_TOL * abs(float(c["loss"]))
    # Tolerance floor: d/d(att_dst) is a cancellation residue (softmax is shift-invariant per target, so it
    # is nonzero only through LeakyReLU kinks) and sits ~1e-7 below the other gradients; a relative test on
    # such a tensor compares rounding noise.  Every tensor is therefore held to GRAD_TOL relative to
    # max(its own scale, 1e-3 x the largest parameter-gradient norm of the model).
    floor = 1e-3 * max(c["grad_norms"].values())
    for k, p in model.named_parameters():
        gn = c["grad_norms"][k]
        assert abs(float(p.grad.norm()) - gn) <= GRAD_TOL * max(gn, floor), f"{k}: |grad| {float(p.grad.norm())} vs {gn}"
        head = p.grad.reshape(-1)[:8].cpu()
        assert float((head - c["grad_heads"][k]).abs().max()) <= GRAD_TOL * max(float(c["grad_heads"][k].abs().max()), gn / p.numel() ** 0.5, floor), k
        if "grads" in c:
            ref = c["grads"][k]
            assert float((p.grad.cpu() - ref).abs().max()) <= GRAD_TOL * max(float(ref.abs().max()), floor), k


@pytest.mark.parametrize("cluster,threads", [(1, 256), (2, 256), (4, 256), (8, 256), (0, 256), (8, 255), (2, 255), (0, 255)])
@pytest.mark.parametrize("graph,B,blocks", [("tiny", 5, 2), ("directed", 3, 3), ("ctown", 4, 4), ("ctown", 40, 2)])
def test_resident_kernels_equal_layer_kernels(graph, B, blocks, cluster, threads, dev, G):
    """The snapshot-resident cluster kernels (small batches) against the layer-by-layer kernels on the same
    weights and inputs: forward (training and inference) and every parameter gradient, for each cluster size.
    `directed` has in-degrees above 8 (chunked online softmax) and an asymmetric transposed CSR; `tiny` leaves
    CTAs of the cluster without rows."""
    from gnn_pressure_estimation_b200 import _lib
    lib = _lib.load()
    ei, names = GRAPHS[graph]() if graph != "directed" else (random_directed_graph(50, 420, 3).numpy(), list(range(50)))
    ei = torch.as_tensor(ei)
    N = len(names)
    ref = O.make_oracle(blocks, 32, seed=11)
    x, y, mask = O.synthetic_snapshots(N, B, seed=5)
    eib = O.collate_edge_index(ei, N, B).to(dev)
    results = []
    prev_c = lib.gatres_set_resident_cluster(cluster)
    prev_t = lib.gatres_set_resident_threads(threads)
    prev = lib.gatres_set_resident_max_batch(-1)
    try:
        for max_b in (0, 256):
            lib.gatres_set_resident_max_batch(max_b)
            model = G.GATResMeanConv(num_blocks=blocks, nc=32)
            model.load_state_dict(ref.state_dict())
            model = model.to(dev)
            batch = torch.arange(B, device=dev).repeat_interleave(N)
            out = model(x.to(dev), eib, batch, None)
            torch.nn.functional.mse_loss(out[mask.to(dev)], y.to(dev)[mask.to(dev)]).backward()
            with torch.no_grad():
                out_inf = model(x.to(dev), eib, batch, None)
            assert torch.equal(out.detach(), out_inf)
            results.append((out.detach(), {k: p.grad.clone() for k, p in model.named_parameters()}))
    finally:
        lib.gatres_set_resident_max_batch(prev)
        lib.gatres_set_resident_cluster(prev_c)
        lib.gatres_set_resident_threads(prev_t)
    (out_l, g_l), (out_r, g_r) = results
    assert_close(out_r, out_l, 1e-5, "resident forward vs layer kernels")
    floor = 1e-3 * max(float(g.norm()) for g in g_l.values())
    for k in g_l:
        err = float((g_r[k] - g_l[k]).abs().max())
        assert err <= 1e-4 * max(float(g_l[k].abs().max()), floor), f"{k}: {err}"
    # and against the CPU oracle
    out_ref, _, grads_ref = O.train_step_loss_and_grads(ref, x, y, mask, O.collate_edge_index(ei, N, B))
    assert_close(out_r, out_ref, FWD_TOL, "resident forward vs oracle")
    floor = 1e-3 * max(float(g.norm()) for g in grads_ref.values())
    for k, g in grads_ref.items():
        assert float((g_r[k].cpu() - g).abs().max()) <= GRAD_TOL * max(float(g.abs().max()), floor), k


@pytest.mark.parametrize("graph,B,blocks", [("ctown", 3, 10), ("tiny", 4, 3), ("ctown", 2, 2)])
def test_sibling_gat_model_matches_oracle(graph, B, blocks, dev, G):
    """the reference's plain `GAT` baseline (GraphModels.py:210-230, `select_model("gat")`) on the
    same fused kernels; its 1 -> 2x32 first layer and 64 -> 1x1 last layer run zero-padded to built shapes."""
    import argparse
    from gnn_pressure_estimation_b200 import ConfigModels as CM
    ei_np, names = GRAPHS[graph]()
    ei, N = torch.from_numpy(ei_np), len(names)
    ref = O.make_gat_oracle(blocks, 32, seed=2)
    if blocks == 10:
        args, model = CM.select_model(argparse.Namespace(model="gat"))
        assert (args.criterion, args.norm_type) == ("mse", "znorm") and model.name == "GAT_10b_32c_2h_10b_32c"
    else:
        model = G.GAT(num_blocks=blocks, nc=32)
    assert list(model.state_dict()) == list(ref.state_dict())
    model.load_state_dict(ref.state_dict())
    model = model.to(dev)
    x, y, mask = O.synthetic_snapshots(N, B, seed=21)
    eib = O.collate_edge_index(ei, N, B)
    out_ref, loss_ref, grads_ref = O.train_step_loss_and_grads(ref, x, y, mask, eib)
    out = model(x.to(dev), eib.to(dev), None, None)
    loss = torch.nn.functional.mse_loss(out[mask.to(dev)], y.to(dev)[mask.to(dev)])
    loss.backward()
    assert out.shape == out_ref.shape
    assert_close(out, out_ref, FWD_TOL, "GAT forward")
    floor = 1e-3 * max(float(g.norm()) for g in grads_ref.values())
    for k, p in model.named_parameters():
        g = grads_ref[k]
        assert p.grad.shape == g.shape, k
        assert float((p.grad.cpu() - g).abs().max()) <= GRAD_TOL * max(float(g.abs().max()), floor), k


def test_mixed_topology_batch_runs_in_general_mode(dev, G, kernel_variant):
    """WDNDataset accepts several networks and shuffle=True can collate different graphs into one batch
    (utils/DataLoader.py:120-129): such an edge_index is not B copies of a template, the model then treats the whole
    collated graph as one snapshot (B = 1, N = sum of the node counts) on the same kernels."""
    ei_a, names_a = GRAPHS["tiny"]()
    ei_b, names_b = GRAPHS["ctown"]()
    na, nb_ = len(names_a), len(names_b)
    ei = torch.cat([torch.from_numpy(ei_a), torch.from_numpy(ei_b) + na, torch.from_numpy(ei_a) + na + nb_], dim=1)
    M = 2 * na + nb_
    ref = O.make_oracle(3, 32, seed=4)
    model = G.GATResMeanConv(num_blocks=3, nc=32)
    model.load_state_dict(ref.state_dict())
    model = model.to(dev)
    g = torch.Generator().manual_seed(8)
    x = torch.randn(M, 1, generator=g)
    y = torch.randn(M, 1, generator=g)
    out_ref = ref(x, ei, None, None)
    torch.nn.functional.mse_loss(out_ref, y).backward()
    out = model(x.to(dev), ei.to(dev), None, None)
    torch.nn.functional.mse_loss(out, y.to(dev)).backward()
    assert_close(out, out_ref, FWD_TOL, "mixed-topology forward")
    gref = {k: q.grad for k, q in ref.named_parameters()}
    floor = 1e-3 * max(float(q.norm()) for q in gref.values())
    for k, p in model.named_parameters():
        assert float((p.grad.cpu() - gref[k]).abs().max()) <= GRAD_TOL * max(float(gref[k].abs().max()), floor), k


def test_model_inference_equals_training_forward_and_batch_hint(dev, G):
    c = load_case("ctown_small_15b_32c_B8")
    model, _ = _cuda_model_from_case(c, G, dev)
    eib = O.collate_edge_index(c["edge_index"], c["N"], c["B"]).to(dev)
    x = c["x"].to(dev)
    out_train = model(x, eib)
    batch = torch.arange(c["B"], device=dev).repeat_interleave(c["N"])
    with torch.no_grad():
        out_inf = model(x, eib, batch, None)
    assert torch.equal(out_train.detach(), out_inf)        # same kernels, rolling buffers instead of saved ones
    assert_close(out_inf, c["out"], FWD_TOL, "inference forward")


def test_large_batch_properties(dev, G):
    """BASELINE size (B=1024 snapshots): size-independent properties instead of the oracle."""
    c = load_case("ctown_small_15b_32c_B8")
    model, _ = _cuda_model_from_case(c, G, dev)
    N, B = c["N"], 1024
    ei = c["edge_index"]
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B * N, 1, generator=g).to(dev)
    with torch.no_grad():
        full = model(x, O.collate_edge_index(ei, N, B).to(dev))
        half = O.collate_edge_index(ei, N, B // 2).to(dev)
        lo, hi = model(x[: B // 2 * N], half), model(x[B // 2 * N:], half)
        assert torch.equal(full, torch.cat([lo, hi]))      # snapshots are independent, kernels deterministic
        perm = torch.randperm(B, generator=g).to(dev)
        xp = x.view(B, N)[perm].reshape(-1, 1)
        assert torch.equal(model(xp, O.collate_edge_index(ei, N, B).to(dev)).view(B, N), full.view(B, N)[perm])
        # the first 8 snapshots against the oracle-pinned golden inputs
        x8 = c["x"].to(dev)
        assert_close(model(torch.cat([x8, x[8 * N:]]), O.collate_edge_index(ei, N, B).to(dev))[: 8 * N], c["out"],
                     FWD_TOL, "B=1024 prefix")


def test_gradient_is_mean_over_shards(dev, G):
    """DP contract: grads of the full batch == average of the shard grads."""
    c = load_case("ctown_small_15b_32c_B8")
    N = c["N"]
    grads = []
    for lo, hi in ((0, 8), (0, 4), (4, 8)):
        model, _ = _cuda_model_from_case(c, G, dev)
        sl = slice(lo * N, hi * N)
        eib = O.collate_edge_index(c["edge_index"], N, hi - lo).to(dev)
        mask = c["mask"][sl].to(dev)
        out = model(c["x"][sl].to(dev), eib)
        torch.nn.functional.mse_loss(out[mask], c["y"][sl].to(dev)[mask]).backward()
        grads.append(torch.cat([p.grad.reshape(-1) for p in model.ordered_parameters()]))
    assert_close(0.5 * (grads[1] + grads[2]), grads[0], 1e-4, "shard-mean gradient")


def test_backward_in_block_ranges_equals_one_call(dev, G):
    """gatres_backward_range (the pieces the DP all-reduce overlaps with) == gatres_backward, and each range's
    slice of the flat gradient buffer is final when the range returns."""
    import ctypes as C
    from gnn_pressure_estimation_b200 import _lib, dp
    from gnn_pressure_estimation_b200.train_step import TrainStep
    c = load_case("ctown_small_15b_32c_B8")
    model, _ = _cuda_model_from_case(c, G, dev)
    N, B = c["N"], c["B"]
    topo = model.set_topology(c["edge_index"].to(dev), N)
    ts = TrainStep(model, topo, B, mask_count_per_snapshot=int(N * 0.95), use_graph=False)
    ts.load_inputs(c["y"].to(dev), c["y"].to(dev), c["mask"].to(dev))
    lib = _lib.load()
    s, d = _lib.stream(), C.byref(ts.desc)
    p = _lib.ptr
    _lib.call("gatres_apply_mask", p(ts.x), p(ts.mask), p(ts.xm), ts.M, s)
    _lib.call("gatres_forward", d, p(ts.flat), p(ts.xm), p(ts.out), p(ts.saved), p(ts.scratch), s)
    _lib.call("gatres_masked_mse", p(ts.out), p(ts.y), p(ts.mask), ts.M, ts.count, p(ts.d_out), p(ts.loss),
              p(ts._loss_part), s)
    _lib.call("gatres_backward", d, p(ts.flat), p(ts.xm), p(ts.saved), p(ts.d_out), None, p(ts.grads), p(ts.scratch), s)
    whole = ts.grads.clone()
    off = lambda k: int(lib.gatres_param_offset_of_block(ts.nb, ts.nc, k))
    assert off(0) == 2 * ts.nc and off(ts.nb) == ts.P - ts.nc - 1 and off(-1) == 0
    for buckets in (2, 3, 15):
        ts.grads.fill_(float("nan"))                       # the head range must zero the buffer itself
        done_from = ts.P
        for k_hi, k_lo in dp.bucket_ranges(ts.nb, buckets):
            _lib.call("gatres_backward_range", d, p(ts.flat), p(ts.xm), p(ts.saved), p(ts.d_out), None, p(ts.grads),
                      p(ts.scratch), k_hi, k_lo, s)
            lo, hi = dp.bucket_slice(ts.nb, ts.P, k_hi, k_lo, off)
            assert hi == done_from
            done_from = lo
            # atomic accumulation order differs between runs -> tolerance, not equality
            assert_close(ts.grads[lo:], whole[lo:], 1e-4, f"{buckets} buckets, final slice after blocks {k_hi}..{k_lo}")
        assert done_from == 0
    with pytest.raises(_lib.GatresError):
        _lib.call("gatres_backward_range", d, p(ts.flat), p(ts.xm), p(ts.saved), p(ts.d_out), None, p(ts.grads),
                  p(ts.scratch), 3, 5, s)


# --------------------------------------------------------------- train step
@pytest.mark.parametrize("use_graph", [False, True])
def test_train_step_matches_oracle_adam(use_graph, dev, G):
    from gnn_pressure_estimation_b200.train_step import TrainStep
    c = load_case("ctown_small_15b_32c_B8")
    model, ref = _cuda_model_from_case(c, G, dev)
    N, B = c["N"], c["B"]
    topo = model.set_topology(c["edge_index"].to(dev), N)
    ts = TrainStep(model, topo, B, mask_count_per_snapshot=int(N * 0.95), use_graph=use_graph,
                   deterministic=not use_graph)
    opt = torch.optim.Adam(ref.parameters(), lr=5e-4, weight_decay=6e-6)
    eib = O.collate_edge_index(c["edge_index"], N, B)
    steps = 3
    if use_graph:
        ts.capture()
    for s in range(steps):
        x, y, mask = O.synthetic_snapshots(N, B, seed=100 + s)
        opt.zero_grad()
        out = ref(x, eib)                         # oracle: x already has masked nodes zeroed
        loss_ref = torch.nn.functional.mse_loss(out[mask], y[mask])
        loss_ref.backward()
        opt.step()
        loss = ts.step(y.to(dev), y.to(dev), mask.to(dev))      # product masks on device (x = y unmasked)
        assert abs(float(loss) - float(loss_ref)) <= 2e-4 * abs(float(loss_ref)), f"step {s}"
    flat_ref = torch.cat([p.detach().reshape(-1) for p in
                          [ref.lin0.weight, ref.lin0.bias] +
                          [t for b in ref.blocks for t in (b.conv1.lin_src.weight, b.conv1.att_src, b.conv1.att_dst,
                                                           b.conv1.bias, b.conv2.lin_src.weight, b.conv2.att_src,
                                                           b.conv2.att_dst, b.conv2.bias)] +
                          [ref.lin1.weight, ref.lin1.bias]])
    # Adam's first steps move every weight by ~lr * sign(g).  d/d(att_dst) is a cancellation residue at
    # rounding-noise level (softmax is shift-invariant per target), so its sign — and hence its Adam update —
    # is noise in ANY fp32 implementation, the reference's included: those entries are only bounded by
    # 2*lr*steps; everything else must agree tightly.
    diff = (ts.flat.cpu() - flat_ref).abs()
    lr = 5e-4
    noise = torch.zeros_like(diff, dtype=torch.bool)
    off = 0
    for p, name in zip(model.ordered_parameters(), ["lin0.w", "lin0.b"] + ["W", "att_src", "att_dst", "bias"] * 2 * 15 + ["lin1.w", "lin1.b"]):
        if name == "att_dst":
            noise[off:off + p.numel()] = True
        off += p.numel()
    assert float(diff.max()) <= 2.0 * lr * steps + 1e-7
    assert float((diff[~noise] > 0.05 * lr * steps).float().mean()) < 1e-3, "parameters after Adam steps"
    assert int(ts.step_count.item()) == steps
