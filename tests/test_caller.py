"""Caller-side pieces next to the hot path: masks, (de)normalisation, the seven metrics.

CPU part: the oracle restatement against vectors produced by the REAL reference functions
(tests/golden/caller_ref.npz, generator tests/golden/make_golden_caller.py).
GPU part: the CUDA kernels (through the C ABI) against the oracle and the same vectors."""
import os

import numpy as np
import pytest
import torch

from oracle import caller_oracle as CO
from oracle import gatres_oracle as O

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "caller_ref.npz"))
CASES = ["znorm_typical", "minmax_typical", "raw_small_targets", "raw_offset", "perfect_prediction"]
NORM = {0: None, 1: "znorm", 2: "minmax"}


def _norm_kwargs(case):
    code, mean, std, mn, mx = GOLD[f"{case}/norm"]
    return dict(norm_type=NORM[int(code)], mean=float(mean), std=float(std), min=float(mn), max=float(mx))


def test_metric_names_and_order():
    assert tuple(GOLD["metric_names"]) == CO.METRIC_NAMES


@pytest.mark.parametrize("case", CASES)
def test_oracle_metrics_equal_reference(case):
    kw = _norm_kwargs(case)
    p = CO.descale(torch.from_numpy(GOLD[f"{case}/pred"]), **kw)
    t = CO.descale(torch.from_numpy(GOLD[f"{case}/true"]), **kw)
    assert np.array_equal(p.numpy(), GOLD[f"{case}/pred_descaled"]) and np.array_equal(t.numpy(), GOLD[f"{case}/true_descaled"])
    got = CO.metrics(p, t)
    for name, ref in zip(CO.METRIC_NAMES, GOLD[f"{case}/metrics"]):
        assert float(got[name]) == pytest.approx(float(ref), rel=1e-6, abs=1e-7), name


def test_oracle_scale_equals_reference():
    x = GOLD["scale/x"]
    assert np.array_equal(CO.scale(x, "znorm", mean=57.3, std=21.9), GOLD["scale/znorm"])
    assert np.array_equal(CO.scale(x, "minmax", min=-3.5, max=140.25), GOLD["scale/minmax"])


def test_oracle_masks_equal_reference_bit_exact():
    np.random.seed(1234)
    m = CO.generate_batch_mask([388] * 6, 0.95)
    assert np.array_equal(m, GOLD["mask/ctown_B6"]) and m.reshape(6, 388).sum(1).tolist() == [368] * 6
    # the seeded-RandomState helper the parity tests use draws the same stream
    assert np.array_equal(O.generate_batch_mask(388, 6, 0.95, seed=1234), GOLD["mask/ctown_B6"])
    np.random.seed(7)
    r = CO.generate_batch_mask([7, 40, 388], 0.6, required_idx=[0, 3])
    assert np.array_equal(r, GOLD["mask/ragged_required"])
    assert r[0] and r[3] and r[7] and r[10] and r[47] and r[50]           # required nodes of every snapshot


def test_device_mask_key_function_matches_its_restatement():
    """host copy of the kernel's key function (C ABI, no GPU needed) vs the numpy restatement"""
    from gnn_pressure_estimation_b200 import _lib
    lib = _lib.load()
    rows = np.array([0, 1, 2, 387, 388, 12415, 2 ** 31 - 1], dtype=np.int64)
    for seed, step in [(0, 0), (1234, 7), (2 ** 63 + 5, 2 ** 40 + 3)]:
        got = np.array([lib.gatres_mask_key(seed, step, int(r)) for r in rows], dtype=np.uint32)
        assert np.array_equal(got, CO.device_mask_keys(seed, step, rows))
    m = CO.device_mask_reference(5, 3, 4, 388, 368).reshape(4, 388)
    assert m.sum(1).tolist() == [368] * 4 and not np.array_equal(m[0], m[1])


def test_numpy_compatible_mask_mode_equals_reference():
    """the product's host-side NumPy mode draws the reference's mask stream bit for bit"""
    from gnn_pressure_estimation_b200.metrics import numpy_batch_mask
    np.random.seed(1234)
    assert np.array_equal(numpy_batch_mask([388] * 6, 0.95), GOLD["mask/ctown_B6"])
    np.random.seed(7)
    assert np.array_equal(numpy_batch_mask([7, 40, 388], 0.6, [0, 3]), GOLD["mask/ragged_required"])


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_metrics_equal_reference(case):
    from gnn_pressure_estimation_b200.metrics import METRIC_NAMES, MaskedMetrics
    dev = torch.device("cuda:0")
    assert METRIC_NAMES == CO.METRIC_NAMES
    kw = _norm_kwargs(case)
    p, t = torch.from_numpy(GOLD[f"{case}/pred"]).to(dev), torch.from_numpy(GOLD[f"{case}/true"]).to(dev)
    mm = MaskedMetrics(dev, prefix="m", **kw)
    got = mm.update(p, t, None).cpu().double().numpy()
    ref = GOLD[f"{case}/metrics"]
    # reference = fp32 torch reductions, ours = fp64 accumulation of the same fp32 terms
    for i, name in enumerate(METRIC_NAMES):
        assert got[i] == pytest.approx(ref[i], rel=2e-5, abs=1e-6), name
    assert got[7] == p.numel()
    assert list(mm.as_dict()) == [f"m_{k}" for k in METRIC_NAMES]
    # masked form: scatter the entries into a larger batch, metrics over mask only (train.py:177-178)
    g = torch.Generator().manual_seed(1)
    M = 3 * p.numel()
    sel = torch.randperm(M, generator=g)[: p.numel()].sort().values.to(dev)
    out_full, y_full = torch.randn(M, generator=g).to(dev) * 50, torch.randn(M, generator=g).to(dev) * 50
    mask = torch.zeros(M, dtype=torch.bool, device=dev)
    out_full[sel], y_full[sel], mask[sel] = p.reshape(-1), t.reshape(-1), True
    got_m = mm.update(out_full, y_full, mask).cpu().double().numpy()
    assert np.allclose(got_m, got, rtol=1e-6, atol=1e-7, equal_nan=True)


@pytest.mark.gpu
def test_cuda_metrics_large_and_degenerate():
    from gnn_pressure_estimation_b200.metrics import MaskedMetrics
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    t = torch.randn(1024 * 388, generator=g)                       # BASELINE size: 1024 snapshots
    p = t + 0.1 * torch.randn(t.shape, generator=g)
    mask = torch.from_numpy(O.generate_batch_mask(388, 1024, 0.95, seed=3))
    ref = CO.metrics(CO.descale(p[mask].double(), "znorm", mean=57.3, std=21.9), CO.descale(t[mask].double(), "znorm", mean=57.3, std=21.9))
    mm = MaskedMetrics(dev, "znorm", mean=57.3, std=21.9)
    got = mm.update(p.to(dev), t.to(dev), mask.to(dev)).cpu().double()
    for i, name in enumerate(CO.METRIC_NAMES):
        assert float(got[i]) == pytest.approx(float(ref[name]), rel=1e-5, abs=1e-6), name
    assert int(got[7]) == 1024 * 368
    # nothing above the 0.01 threshold -> the reference's mean over an empty selection is nan; constant target -> nan corr
    z = torch.zeros(100, device=dev)
    v = MaskedMetrics(dev, None).update(z + 0.001, z, None).cpu()
    assert torch.isnan(v[0]) and torch.isnan(v[2]) and float(v[4]) == pytest.approx(0.001)


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,rate", [(6, 388, 0.95), (3, 7, 0.6), (2, 100000, 0.95), (1, 1000, 0.001), (5, 257, 1.0)])
def test_cuda_mask_is_exact_count_and_bit_exact(B, N, rate):
    from gnn_pressure_estimation_b200.metrics import generate_batch_mask, mask_count
    dev = torch.device("cuda:0")
    count = mask_count(N, rate)
    for seed, step in [(1234, 0), (1234, 1), (99, 12345)]:
        m = generate_batch_mask(B, N, rate, seed, step, device=dev).cpu().numpy().astype(bool)
        assert m.reshape(B, N).sum(1).tolist() == [count] * B                       # auxil.py:160 assertion
        assert np.array_equal(m, CO.device_mask_reference(seed, step, B, N, count))
    # required nodes (evaluation.py:288-291 sensors) are in every snapshot's mask, the rest is drawn around them
    if N >= 40 and count >= 4:
        from gnn_pressure_estimation_b200.metrics import required_flags
        req = [0, 3, N - 1]
        m = generate_batch_mask(B, N, rate, 5, 2, required=required_flags(N, req, dev), device=dev).cpu().numpy().astype(bool)
        assert m.reshape(B, N).sum(1).tolist() == [count] * B and m.reshape(B, N)[:, req].all()
        assert np.array_equal(m, CO.device_mask_reference(5, 2, B, N, count, req))
    # step_dev is added to step on the device (CUDA-graph replay draws a fresh mask)
    sd = torch.tensor([5], dtype=torch.int32, device=dev)
    a = generate_batch_mask(B, N, rate, 7, 10, step_dev=sd, device=dev)
    b = generate_batch_mask(B, N, rate, 7, 15, device=dev)
    assert torch.equal(a, b)


@pytest.mark.gpu
def test_cuda_mask_is_uniform():
    """every node is masked with probability count/N; pairs of nodes are (almost) independent"""
    from gnn_pressure_estimation_b200.metrics import generate_batch_mask
    dev = torch.device("cuda:0")
    B, N, rate = 4096, 97, 0.5
    m = generate_batch_mask(B, N, rate, seed=2024, step=0, device=dev).view(B, N).float()
    count = int(N * rate)
    freq = m.mean(0).cpu().numpy()
    sigma = (count / N * (1 - count / N) / B) ** 0.5
    assert np.abs(freq - count / N).max() < 5 * sigma
    cov = (m.T @ m / B).cpu().numpy()
    expect = count * (count - 1) / (N * (N - 1))                                    # P(both masked), without replacement
    off = cov[~np.eye(N, dtype=bool)]
    assert np.abs(off - expect).max() < 6 * (expect * (1 - expect) / B) ** 0.5


@pytest.mark.gpu
def test_train_step_with_device_mask_and_metrics():
    """TrainStep drawing its masks on the device == TrainStep fed the same masks from the host; metrics per step"""
    import gnn_pressure_estimation_b200.GraphModels as G
    from gnn_pressure_estimation_b200.metrics import MaskedMetrics
    from gnn_pressure_estimation_b200.train_step import TrainStep
    from helpers import load_case
    dev = torch.device("cuda:0")
    c = load_case("ctown_small_15b_32c_B8")
    N, B, count = c["N"], c["B"], int(c["N"] * 0.95)
    ref = O.make_oracle(3, 32, seed=0)
    runs = []
    for device_mask in (True, False):
        model = G.GATResMeanConv(num_blocks=3, nc=32)
        model.load_state_dict(ref.state_dict())
        model = model.to(dev)
        topo = model.set_topology(c["edge_index"].to(dev), N)
        mm = MaskedMetrics(dev, "znorm", mean=57.3, std=21.9)
        ts = TrainStep(model, topo, B, count, use_graph=True, device_mask_seed=42 if device_mask else None, metrics=mm)
        ts.capture()
        losses = []
        for s in range(3):
            _, y, _ = O.synthetic_snapshots(N, B, seed=300 + s)
            host_mask = torch.from_numpy(CO.device_mask_reference(42, s, B, N, count))
            loss = ts.step(y.to(dev), y.to(dev), None if device_mask else host_mask.to(dev))
            losses.append(float(loss))
            if device_mask:
                assert torch.equal(ts.mask.cpu().bool(), host_mask), f"step {s}"
            m = mm.as_dict()
            sel = host_mask
            want = CO.metrics(CO.descale(ts.out.cpu()[sel].double(), "znorm", mean=57.3, std=21.9),
                              CO.descale(y.reshape(-1)[sel].double(), "znorm", mean=57.3, std=21.9))
            for k in CO.METRIC_NAMES:
                assert m[f"tr_{k}"] == pytest.approx(float(want[k]), rel=1e-4, abs=1e-6), (s, k)
        runs.append((losses, ts.flat.clone()))
    assert runs[0][0] == pytest.approx(runs[1][0], rel=1e-5)
    # (d/d att_dst is rounding noise whose sign Adam amplifies to +-lr per step: see test_train_step_matches_oracle_adam)
    assert float((runs[0][1] - runs[1][1]).abs().max()) <= 2 * 5e-4 * 3 + 1e-6


def test_timer_arithmetic_equals_reference():
    """compute_time / compute_throughput (utils/timer.py:43-66, incl. its batch-count quirk) vs the real reference"""
    from gnn_pressure_estimation_b200 import evaluation as E
    t, g = GOLD["timer/timings"].tolist(), GOLD["timer/num_graphs"].tolist()
    ref_time, ref_thr = GOLD["timer/time_throughput"]
    assert E.compute_time(t, g, 107) == pytest.approx(ref_time, rel=1e-12)
    assert E.compute_throughput(t, g, 107) == pytest.approx(ref_thr, rel=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("required,use_same_mask", [((), False), ((3, 17, 200), False), ((), True)])
def test_evaluation_epoch_matches_oracle(required, use_same_mask):
    """test_one_epoch on the CUDA kernels vs the oracle loop with the same NumPy mask stream (evaluation.py:300-351)"""
    import gnn_pressure_estimation_b200.GraphModels as G
    from gnn_pressure_estimation_b200 import evaluation as E
    from helpers import load_case
    dev = torch.device("cuda:0")
    c = load_case("ctown_small_15b_32c_B8")
    N, ei = c["N"], c["edge_index"]
    ref = O.make_oracle(3, 32, seed=0)
    model = G.GATResMeanConv(num_blocks=3, nc=32)
    model.load_state_dict(ref.state_dict())
    model = model.to(dev)
    g = torch.Generator().manual_seed(9)
    S, bs = 20, 8 if not use_same_mask else 10                       # 8 + 8 + 4: a smaller last batch, as without drop_last
    snaps = torch.randn(S, N, generator=g)
    kw = dict(norm_type="znorm", mean=57.3, std=21.9)
    np.random.seed(77)
    loss_ref, m_ref = CO.test_one_epoch_oracle(ref, snaps, ei, bs, 0.95, required_idx=required, use_same_mask=use_same_mask, **kw)
    np.random.seed(77)
    loss, m = E.test_one_epoch(model, snaps.to(dev), ei, bs, 0.95, required_idx=required, use_same_mask=use_same_mask,
                               gpu_warmup_times=2, mask_source="numpy", **kw)
    post = "_sensor" if required else ""
    assert loss == pytest.approx(loss_ref, rel=2e-4)
    for k in CO.METRIC_NAMES:
        assert m[f"test_{k}{post}"] == pytest.approx(m_ref[k], rel=2e-4, abs=1e-6), k
    assert m[f"test_time{post}"] > 0 and m[f"test_throughput{post}"] > 0 and m[f"test_snapshots_per_s{post}"] > 0
    # device masks: same contract, different stream -> statistically the same figures
    loss_d, m_d = E.test_one_epoch(model, snaps.to(dev), ei, bs, 0.95, required_idx=required, use_same_mask=use_same_mask,
                                   gpu_warmup_times=0, mask_source="device", seed=5, **kw)
    assert loss_d == pytest.approx(loss_ref, rel=0.1) and m_d[f"test_rmse{post}"] == pytest.approx(m_ref["rmse"], rel=0.1)


@pytest.mark.parametrize("case", CASES)
def test_descale_affine_equals_reference(case):
    """every norm of utils/auxil.py:42-64 is v * scale + shift with the product's (scale, shift)"""
    from gnn_pressure_estimation_b200.metrics import descale_affine
    kw = _norm_kwargs(case)
    scale, shift = descale_affine(kw["norm_type"], mean=kw["mean"], std=kw["std"], min=kw["min"], max=kw["max"])
    p = torch.from_numpy(GOLD[f"{case}/pred"])
    assert np.array_equal((p * scale + shift).numpy(), GOLD[f"{case}/pred_descaled"])
    with pytest.raises(ValueError):
        descale_affine("znorm")
