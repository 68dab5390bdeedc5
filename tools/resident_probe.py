"""Time the whole-model forward / backward C calls (gatres_forward, gatres_backward) on one GPU for a sweep of
batch sizes, resident-kernel cluster sizes and the layer-by-layer path.  CUDA events around graph replays.

    python tools/resident_probe.py --batches 8,32,64,128,256 --clusters 0,2,4,8 --out gpurun_out/probe.json
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gnn_pressure_estimation_b200 import _lib, topology  # noqa: E402
from gnn_pressure_estimation_b200.GraphModels import GATResMeanConv  # noqa: E402
from gnn_pressure_estimation_b200.graph import Topology  # noqa: E402
from gnn_pressure_estimation_b200.train_step import TrainStep  # noqa: E402


def time_graph(fn, reps=50):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(5):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


FWD_PHASES = ["proj1", "sync", "agg1", "proj2", "sync", "agg2", "sync", "mean+wait"]
BWD_PHASES = ["mean_bwd+p1(conv2)", "sync", "p2(conv2)", "wgrad2+dgrad2", "p1(conv1)", "sync", "p2(conv1)",
              "wgrad1+dgrad1+vec", "sync"]


def phases(lib, fwd, bwd, blocks, dev, only=None):
    """mean duration (us) of each phase of a block, over blocks 1.. and all CTAs, from %globaltimer stamps"""
    slots, ctas = 1 + 9 * blocks + 2, 4096
    buf = torch.zeros(ctas * slots, dtype=torch.int64, device=dev)
    out = {}
    for name, fn, per, labels in (("fwd", fwd, 8, FWD_PHASES), ("bwd", bwd, 9, BWD_PHASES)):
        if only is not None and name not in only:
            continue
        buf.zero_()
        lib.gatres_set_resident_profile(buf.data_ptr(), slots)
        fn()
        fn()
        torch.cuda.synchronize()
        lib.gatres_set_resident_profile(None, 0)
        t = buf.view(ctas, slots).cpu()
        t = t[t[:, 0] > 0].double()
        body = t[:, 1:1 + per * blocks].view(-1, blocks, per)
        nxt = torch.cat([body[:, 1:, 0], t[:, 1 + per * blocks:2 + per * blocks] if name == "bwd" else body[:, -1:, -1]], dim=1)
        ends = torch.cat([body[:, :, 1:], nxt.unsqueeze(-1)], dim=2)
        dur = (ends - body)[:, 1:-1].mean(dim=(0, 1)) / 1e3
        out[name] = {f"{i}:{l}": round(float(d), 3) for i, (l, d) in enumerate(zip(labels, dur))}
        out[name]["kernel_us"] = round(float((t[:, :1 + per * blocks].max() - t[:, 0].min()) / 1e3), 1)
        out[name]["ctas"] = int(t.shape[0])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="32")
    ap.add_argument("--clusters", default="0")
    ap.add_argument("--threads", default="256", help="256 = tensor-core contractions, 255 = FFMA contractions")
    ap.add_argument("--blocks", type=int, default=15)
    ap.add_argument("--layer", action="store_true", help="also time the layer-by-layer kernels")
    ap.add_argument("--phases", action="store_true", help="per-phase timestamps of the resident kernels")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = _lib.load()
    ei_np, names = topology.reference_edge_index(topology.ctown_shaped())
    N = len(names)
    topo = Topology.build(torch.from_numpy(ei_np).to(dev), N)
    torch.manual_seed(0)
    model = GATResMeanConv(num_blocks=a.blocks, nc=32).to(dev)
    rows = []
    for B in [int(b) for b in a.batches.split(",")]:
        ts = TrainStep(model, topo, B, int(N * 0.95), use_graph=False)
        ts.load_inputs(torch.randn(B * N, device=dev), torch.randn(B * N, device=dev),
                       (torch.rand(B * N, device=dev) < 0.95))
        d, p, s = C.byref(ts.desc), _lib.ptr, _lib.stream
        variants = [("resident", int(c), int(t)) for c in a.clusters.split(",") for t in a.threads.split(",")] + \
            ([("layer", 0, 0)] if a.layer else [])
        for kind, cs, thr in variants:
            if thr:
                lib.gatres_set_resident_threads(thr)
            lib.gatres_set_resident_max_batch(1 << 30 if kind == "resident" else 0)
            lib.gatres_set_resident_cluster(cs)
            scratch_inf = torch.empty(int(lib.gatres_scratch_floats(d, 0)), device=dev)
            fwd = lambda: _lib.call("gatres_forward", d, p(ts.flat), p(ts.xm), p(ts.out), p(ts.saved), p(ts.scratch), s())
            inf = lambda: _lib.call("gatres_forward", d, p(ts.flat), p(ts.xm), p(ts.out), None, p(scratch_inf), s())
            bwd = lambda: _lib.call("gatres_backward", d, p(ts.flat), p(ts.xm), p(ts.saved), p(ts.d_out), None,
                                    p(ts.grads), p(ts.scratch), s())
            _lib.call("gatres_apply_mask", p(ts.x), p(ts.mask), p(ts.xm), ts.M, s())
            fwd()
            _lib.call("gatres_masked_mse", p(ts.out), p(ts.y), p(ts.mask), ts.M, ts.count, p(ts.d_out), p(ts.loss),
                      p(ts._loss_part), s())
            row = {"B": B, "kind": kind, "cluster": cs, "threads": thr, "fwd_train_us": time_graph(fwd), "fwd_infer_us": time_graph(inf),
                   "bwd_us": time_graph(bwd)}
            row["step_us"] = time_graph(ts._enqueue_impl)
            if a.phases and kind == "resident":
                row["phases"] = phases(lib, fwd, bwd, a.blocks, dev)
                row["phases_infer"] = phases(lib, inf, bwd, a.blocks, dev, only=("fwd",))
            rows.append(row)
            print(json.dumps(row), flush=True)
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
