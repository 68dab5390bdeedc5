"""Golden vectors for the caller-side pieces, produced by the REAL reference functions
(/root/reference/gnn_pressure_estimation/utils/auxil.py is plain numpy/torch and imports here).

    python tests/golden/make_golden_caller.py        # needs /root/reference; writes tests/golden/caller_ref.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/gnn_pressure_estimation")
import utils.auxil as A  # noqa: E402  (the reference module itself)


def cases():
    g = torch.Generator().manual_seed(2024)
    out = {}
    # (name, y_true scaled, prediction noise, norm)
    t = torch.randn(4 * 368, 1, generator=g)
    out["znorm_typical"] = (t + 0.05 * torch.randn(t.shape, generator=g), t, ("znorm", 57.3, 21.9, None, None))
    t = torch.rand(3 * 368, 1, generator=g)
    out["minmax_typical"] = (t + 0.02 * torch.randn(t.shape, generator=g), t, ("minmax", None, None, -3.5, 140.25))
    t = torch.randn(777, 1, generator=g) * 0.02          # many |y| <= 0.01 (excluded from the relative error), negatives
    out["raw_small_targets"] = (t + 0.01 * torch.randn(t.shape, generator=g), t, (None, None, None, None, None))
    t = torch.randn(5000, 1, generator=g) * 3.0 + 100.0   # mean >> std: cancellation in the correlation / NSE sums
    out["raw_offset"] = (t + 0.3 * torch.randn(t.shape, generator=g), t, (None, None, None, None, None))
    t = torch.randn(64, 1, generator=g)
    out["perfect_prediction"] = (t.clone(), t, ("znorm", 10.0, 2.0, None, None))
    return out


def main():
    blob = {}
    fns = A.get_metric_fn_collection("m")
    for name, (p, t, (norm, mean, std, mn, mx)) in cases().items():
        kw = dict(norm_type=norm, mean=mean, std=std, min=mn, max=mx)
        pd, td = A.descale(scaled_data=p, **kw), A.descale(scaled_data=t, **kw)
        blob[f"{name}/pred"], blob[f"{name}/true"] = p.numpy(), t.numpy()
        blob[f"{name}/norm"] = np.array([{"znorm": 1, "minmax": 2}.get(norm, 0), mean or 0, std or 0, mn or 0, mx or 0], dtype=np.float64)
        blob[f"{name}/pred_descaled"], blob[f"{name}/true_descaled"] = pd.numpy(), td.numpy()
        blob[f"{name}/metrics"] = np.array([float(fn(pd, td)) for fn in fns.values()], dtype=np.float64)
    blob["metric_names"] = np.array([k[2:] for k in fns])
    x = np.linspace(-2.0, 250.0, 17)
    blob["scale/x"] = x
    blob["scale/znorm"] = A.scale(x, "znorm", mean=57.3, std=21.9)
    blob["scale/minmax"] = A.scale(x, "minmax", min=-3.5, max=140.25)
    # masks: the reference consumes the global numpy RNG (train.py:172)
    np.random.seed(1234)
    blob["mask/ctown_B6"] = A.generate_batch_mask(num_nodes=torch.tensor([388] * 6), mask_rate=0.95, required_idx=[])
    np.random.seed(7)
    blob["mask/ragged_required"] = A.generate_batch_mask(num_nodes=[7, 40, 388], mask_rate=0.6, required_idx=[0, 3])
    # Timer arithmetic (utils/timer.py:43-66) without CUDA events: fill the recorded lists directly
    import utils.timer as TM
    tm = object.__new__(TM.Timer)
    tm.timings = [3.25, 3.5, 3.125, 1.75]
    tm.num_graphs = [32, 32, 32, 11]
    blob["timer/timings"], blob["timer/num_graphs"] = np.array(tm.timings), np.array(tm.num_graphs)
    blob["timer/time_throughput"] = np.array([tm.compute_time(107), tm.compute_throughput(107)])
    # EarlyStopping (utils/early_stopping.py:31-78): stop decisions for a validation-loss sequence
    import utils.early_stopping as ES
    seq = [1.0, 0.9, 0.95, 0.91, 0.85, 0.86, 0.87, 0.88, 0.80, float("nan")]
    for pat in (0, 1, 3):
        es = ES.EarlyStopping(mode="min", min_delta=0, patience=pat)
        blob[f"early_stopping/patience{pat}"] = np.array([bool(es.step(torch.tensor(v))) for v in seq])
    blob["early_stopping/seq"] = np.array(seq)
    np.savez_compressed(os.path.join(HERE, "caller_ref.npz"), **blob)
    print({k: v.shape for k, v in blob.items()})


if __name__ == "__main__":
    main()
