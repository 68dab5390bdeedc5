"""CPU checks of the host-side mirror of the reference interface."""
import argparse

import pytest
import torch

import gnn_pressure_estimation_b200.GraphModels as G
from gnn_pressure_estimation_b200 import ConfigModels as CM
from gnn_pressure_estimation_b200 import ops as gops
from oracle import gatres_oracle as O


def test_select_model_contract():
    args = argparse.Namespace(model="gatres_small", model_path="keep/me.pth")
    args, model = CM.select_model(args, None, reset_model_path=True)
    assert isinstance(model, G.GATResMeanConv) and model.num_blocks == 15 and model.nc == 32
    assert (args.criterion, args.norm_type, args.use_data_edge_attrs, args.model_path) == ("mse", "znorm", None, "keep/me.pth")
    assert model.name == "GATResMeanConv_small_znorm_15b_32c"
    args, large = CM.select_model(argparse.Namespace(model="gatres_large"), "variant")
    assert large.num_blocks == 25 and large.nc == 128 and large.name == "variant"
    assert args.model_path.endswith("20235401.pth")
    with pytest.raises(NotImplementedError):
        CM.select_model(argparse.Namespace(model="gin"))


def test_select_model_reaches_the_gat_baseline():
    """ConfigModels.py:96-103: `gat` = 10 GATConv layers, 1 -> 2x32 ... -> 1, PyG-compatible parameter names"""
    args, model = CM.select_model(argparse.Namespace(model="gat"))
    assert isinstance(model, G.GAT) and model.num_blocks == 10 and model.name == "GAT_10b_32c_2h_10b_32c"
    assert (args.criterion, args.norm_type, args.use_data_edge_attrs) == ("mse", "znorm", None)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert shapes["blocks.0.lin_src.weight"] == (64, 1) and shapes["blocks.5.lin_src.weight"] == (64, 64)
    assert shapes["blocks.9.lin_src.weight"] == (1, 64) and shapes["blocks.9.att_src"] == (1, 1, 1) and shapes["blocks.9.bias"] == (1,)
    ref = O.make_gat_oracle(10, 32)
    assert list(model.state_dict()) == list(ref.state_dict())


def test_reference_import_paths():
    from gnn_pressure_estimation.GraphModels import GATResMeanConv, GResBlockMeanConv  # noqa: F401
    from gnn_pressure_estimation.GraphModels import GAT, GATConvNet, GResBlockConv  # noqa: F401  (reference's own import line)
    from gnn_pressure_estimation.ConfigModels import config_gat, select_model  # noqa: F401
    assert GATResMeanConv is G.GATResMeanConv and GAT is G.GAT
    ns = {}
    exec("from gnn_pressure_estimation.GraphModels import *", ns)
    assert {"GAT", "GATConvNet", "GResBlockConv", "GATResMeanConv"} <= set(ns)


def test_every_cli_model_choice_is_selectable():
    """train.py's --model choices must all resolve (gatres_small_tough once raised 'Unknown model')"""
    for which in ("gatres_small", "gatres_large", "gatres_small_tough", "gat"):
        args, model = CM.select_model(argparse.Namespace(model=which))
        assert model is not None and args.criterion == "mse"
    _, tough = CM.select_model(argparse.Namespace(model="gatres_small_tough"))
    assert tough.num_blocks == 15 and tough.nc == 32 and tough.name == "GATResMeanConv_small_tough_znorm_15b_32c"


def test_sibling_models_have_pyg_compatible_parameters():
    """GATConvNet (GraphModels.py:15-46) and GResBlockConv (:548-561): same parameter names / shapes as the oracle
    restatement, so reference checkpoints of those baselines load"""
    net = dict(input_dim=1, hidden_dim=32, heads=2, out_dim=1, num_layers=4)
    m, ref = G.GATConvNet(net), O.GATConvNetOracle(net)
    assert list(m.state_dict()) == list(ref.state_dict())
    assert [tuple(v.shape) for v in m.state_dict().values()] == [tuple(v.shape) for v in ref.state_dict().values()]
    m.load_state_dict(ref.state_dict())
    b, rb = G.GResBlockConv(32, 32, 32), O.OracleGResBlockConv(32, 32, 32)
    assert list(b.state_dict()) == list(rb.state_dict())


def test_state_dict_is_pyg_compatible_both_schemes():
    ref = O.make_oracle(2, 32)
    m = G.GATResMeanConv(num_blocks=2, nc=32)
    assert list(m.state_dict()) == list(ref.state_dict())
    assert [tuple(v.shape) for v in m.state_dict().values()] == [tuple(v.shape) for v in ref.state_dict().values()]
    m.load_state_dict(ref.state_dict())
    new_scheme = {k.replace("lin_src.weight", "lin.weight"): v for k, v in ref.state_dict().items() if "lin_dst" not in k}
    m2 = G.GATResMeanConv(num_blocks=2, nc=32)
    m2.load_state_dict(new_scheme)
    for (k, a), b in zip(m.state_dict().items(), m2.state_dict().values()):
        assert torch.equal(a, b), k
    names = [n for n, _ in m.named_parameters()]
    assert len(names) == 2 + 2 * 8 + 2 and all(("block" in n) == n.startswith("blocks.") for n in names)


def test_flat_parameter_layout_matches_c_abi():
    m = G.GATResMeanConv(num_blocks=3, nc=32)
    flat = m.flat_parameters()
    assert flat.numel() == gops.param_count(3, 32) == sum(p.numel() for p in m.parameters())
    off = 0
    for p in m.ordered_parameters():
        assert p.data_ptr() == flat.data_ptr() + 4 * off        # parameters alias the flat buffer
        assert torch.equal(p.detach().reshape(-1), flat[off:off + p.numel()])
        off += p.numel()
    # in-place updates (optimizers, load_state_dict) keep the aliasing; re-packing is idempotent
    ref = O.make_oracle(3, 32)
    m.load_state_dict(ref.state_dict())
    assert m.flat_parameters().data_ptr() == flat.data_ptr()
    assert torch.equal(flat[:32], ref.lin0.weight.detach().reshape(-1))
    # a dtype/device move breaks aliasing and is repaired on the next call
    m.double()
    with pytest.raises(RuntimeError, match="fp32"):
        m.flat_parameters()
    m.float()
    f2 = m.flat_parameters()
    assert f2.data_ptr() != flat.data_ptr() and torch.equal(f2, flat)


def test_default_init_statistics():
    torch.manual_seed(0)
    m = G.GATResMeanConv(num_blocks=1, nc=32)
    b = m.blocks[0]
    assert float(b.conv1.bias.abs().max()) == 0 and float(b.conv2.bias.abs().max()) == 0
    assert float(b.conv1.lin_src.weight.abs().max()) <= (6 / (64 + 32)) ** 0.5
    assert float(b.conv1.att_src.abs().max()) <= (6 / (2 + 32)) ** 0.5
    assert float(m.lin0.weight.abs().max()) <= 1.0 and float(m.lin1.weight.abs().max()) <= 32 ** -0.5


def test_locality_order_and_permuted_csr():
    """host-side locality plan of the resident kernels (graph.locality_order / permute_csr): a permutation, the same
    rows with the same entry order, and neighbours mostly inside one of eight equal row slices"""
    import numpy as np
    from gnn_pressure_estimation_b200 import graph as Gr, topology as T
    from oracle import topology_oracle as TO
    for wn in (T.tiny_network(), T.ctown_shaped()):
        ei, names = T.reference_edge_index(wn)
        N = len(names)
        perm = Gr.locality_order(ei, N)
        assert sorted(perm.tolist()) == list(range(N))
        assert np.array_equal(perm, Gr.locality_order(ei, N))              # deterministic
        rp, col = TO.csr_by_target(ei, N)
        rpp, colp = Gr.permute_csr(rp.astype(np.int64), col.astype(np.int64), perm)
        assert rpp[-1] == rp[-1] and rpp.dtype == np.int32
        for r in range(N):
            i = perm[r]
            assert np.array_equal(perm[colp[rpp[r]:rpp[r + 1]]], col[rp[i]:rp[i + 1]])
    inv = np.empty(N, dtype=np.int64)
    inv[perm] = np.arange(N)
    R = -(-N // 8)
    before = float((ei[0] // R == ei[1] // R).mean())
    after = float((inv[ei[0]] // R == inv[ei[1]] // R).mean())
    assert before < 0.2 and after > 0.9, (before, after)               # random ids -> 96 % of the neighbours local


def test_degree_sorted_slices_keep_locality_and_remove_padding():
    """graph.degree_sort_slices: rows stay inside their slice (locality share unchanged), the order inside a slice is
    stable by degree, and the warp-level padding (largest degree of 4 / 8 consecutive rows over the mean) shrinks"""
    import numpy as np
    from gnn_pressure_estimation_b200 import graph as Gr, topology as T
    from oracle import topology_oracle as TO
    ei, names = T.reference_edge_index(T.ctown_shaped())
    N = len(names)
    rp, _ = TO.csr_by_target(ei, N)
    deg = (rp[1:] - rp[:-1]).astype(np.int64)
    perm = Gr.locality_order(ei, N)
    sorted_perm = Gr.degree_sort_slices(perm, deg)
    R = -(-N // 8)
    assert sorted(sorted_perm.tolist()) == list(range(N))
    for lo in range(0, N, R):
        assert sorted(sorted_perm[lo:lo + R].tolist()) == sorted(perm[lo:lo + R].tolist())
        assert np.all(np.diff(deg[sorted_perm[lo:lo + R]]) >= 0)

    def padded_work(p, rows_per_warp):
        total = 0
        for lo in range(0, N, R):
            d = deg[p[lo:lo + R]]
            for w in range(0, len(d), rows_per_warp):
                total += int(d[w:w + rows_per_warp].max()) * len(d[w:w + rows_per_warp])
        return total / float(deg.sum())

    for rpw in (4, 8):
        before, after = padded_work(perm, rpw), padded_work(sorted_perm, rpw)
        assert after < before and after < 1.1, (rpw, before, after)
