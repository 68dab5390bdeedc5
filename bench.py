#!/usr/bin/env python
"""Benchmark of the GATRes hot path (contract: see DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
  python bench.py --impl reference [--gpus N] [--steps K] ...    # CPU reference arm (oracle port)

Workload = BASELINE.json configs[1]: gatres_small training on the C-Town-shaped
graph, batch 32 snapshots per GPU (weak scaling over --gpus), synthetic
snapshots, mask_rate 0.95, Adam.  A step = mask + forward + masked MSE +
backward (+ gradient all-reduce) + Adam.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "train snapshots/s, GATRes-small on C-Town"
UNIT = "snapshots/s"
MASK_RATE = 0.95


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="snapshots per GPU per step (configs[1]: 32)")
    ap.add_argument("--model", default="gatres_small", choices=["gatres_small", "gatres_large"])
    ap.add_argument("--graph", default="ctown", choices=["ctown", "scaled"])
    ap.add_argument("--mode", default="train", choices=["train", "infer"])
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-kernel-leg", action="store_true")
    ap.add_argument("--profile-kernels", action="store_true",
                    help="only launch each hot kernel at the HBM-regime batch (for ncu); prints nothing")
    ap.add_argument("--hbm-batch", type=int, default=2048, help="batch of the HBM-regime kernel measurement")
    ap.add_argument("--kernels-json", default=None, help="write the per-kernel table here")
    return ap.parse_args()


def model_cfg(name):
    return (15, 32) if name == "gatres_small" else (25, 128)


def build_graph(kind):
    from gnn_pressure_estimation_b200 import topology as T
    wn = T.ctown_shaped() if kind == "ctown" else T.scaled_wdn()
    ei, names = T.reference_edge_index(wn)
    return torch.from_numpy(ei), len(names)


def workload_name(args):
    g = "C-Town-shaped synthetic graph (N=388, E=858 directed)" if args.graph == "ctown" else \
        "scaled synthetic WDN (N=100000, E=230000 directed)"
    what = "training step (mask+fwd+MSE+bwd+Adam)" if args.mode == "train" else "inference forward"
    return f"{args.model} {what}, {g}, batch {args.batch} snapshots/GPU, mask_rate {MASK_RATE}"


# ----------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock + throttle reasons while the timed region runs (NVML)."""
    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown"}
    NOTE = {0x4: "sw_power_cap"}

    def __init__(self, index):
        self.samples, self.reasons, self._stop = [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in {**self.BAD, **self.NOTE}.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.nv is not None:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv is not None:
            self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------
# CPU reference arm (oracle port)
# ----------------------------------------------------------------------------
def cpu_reference(args, steps, warmup, budget_s):
    """Times the CPU restatement of the reference's PyG op sequence (oracle/ — the
    reference itself cannot run here: torch_geometric is not installable)."""
    from oracle import gatres_oracle as O
    nb, nc = model_cfg(args.model)
    ei, N = build_graph(args.graph)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = O.make_oracle(nb, nc, seed=0)
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=6e-6)

    def one(B, seed):
        x, y, mask = O.synthetic_snapshots(N, B, MASK_RATE, seed=seed)
        eib = O.collate_edge_index(ei, N, B)
        t = time.perf_counter()
        if args.mode == "train":
            opt.zero_grad()
            out = model(x, eib, None, None)
            loss = torch.nn.functional.mse_loss(out[mask], y[mask])
            loss.backward()
            opt.step()
            float(loss)
        else:
            with torch.no_grad():
                model(x, eib, None, None)
        return time.perf_counter() - t

    B = args.batch
    one(min(B, 8), 0)                                   # first-touch / thread-pool spin-up
    probe = one(min(B, 8), 1) / min(B, 8)               # seconds per snapshot
    while B > 1 and probe * B * (steps + warmup) > budget_s:
        B //= 2                                         # bounded sample: fewer snapshots per step
    for w in range(warmup):
        one(B, 10 + w)
    ts = [one(B, 100 + s) for s in range(steps)]
    sec = float(np.sum(ts))
    value = B * steps / sec
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} steps of batch {B} on {cores} host threads (oracle port of the PyG op sequence; "
                      f"median {1e3 * float(np.median(ts)):.1f} ms/step)"}, 1e3 * sec / steps, B


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    cb, ms, B = cpu_reference(args, steps, warmup, budget_s=150.0)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "sample_batch": B, "where": "host CPU"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------
# per-kernel leg (roofline)
# ----------------------------------------------------------------------------
def time_launches(fn, reps, n_sets):
    """average device time of fn(set_index) over `reps` graph-replayed launches (CUDA events on the launch stream)."""
    s = torch.cuda.current_stream()
    for k in range(n_sets):
        fn(k)
    torch.cuda.synchronize()
    if os.environ.get("GATRES_PROFILE_EAGER"):        # under ncu: plain launches, nothing to time
        return 1.0
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for r in range(reps):
            fn(r % n_sets)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    g.replay()
    e1.record(s)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def kernel_table(B, N, topo, nc, hbm_regime, dev):
    """GB/s of each hot kernel of one block on [B*N] rows; algorithmic bytes per node from SURVEY §8d."""
    from gnn_pressure_estimation_b200 import ops as gops
    from gnn_pressure_estimation_b200._lib import call, ptr, stream
    M = B * N
    f = dict(dtype=torch.float32, device=dev)
    n_sets = 2 if hbm_regime else 4
    reps = 20 if hbm_regime else 60
    rows = []

    def rnd(*shape):
        return [torch.randn(*shape, **f) for _ in range(n_sets)]

    for H, K in ((2, nc), (1, 2 * nc)):
        F = H * nc
        x, h = rnd(M, K), rnd(M, F)
        W = torch.randn(F, K, **f) * 0.1
        a_s, a_d, bias = torch.randn(F, **f), torch.randn(F, **f), torch.randn(F, **f)
        ss, sd = rnd(M, H), rnd(M, H)
        out = rnd(M, F)
        m, l = rnd(M, H), rnd(M, H)
        g = rnd(M, F)
        rec, dsd = torch.empty(M, H, 4, **f), torch.empty(M, H, **f)
        dh = torch.empty(M, F, **f)
        dx = torch.empty(M, K, **f)
        S = 0                                   # atomic gradient accumulation, as in the training step
        P = gops.a4(F * K + 3 * F)
        partial = torch.zeros(P, **f)

        def proj(k):
            call("gatres_linear_att_fwd", ptr(x[k]), ptr(W), ptr(a_s), ptr(a_d), ptr(h[k]), ptr(ss[k]), ptr(sd[k]), M, K,
                 H, nc, stream())

        def agg(k):
            call("gatres_gat_agg_fwd", ptr(topo.rowptr), ptr(topo.col), ptr(h[k]), ptr(ss[k]), ptr(sd[k]), ptr(bias),
                 ptr(out[k]), ptr(m[k]), ptr(l[k]), B, N, topo.E1, H, nc, 1, stream())

        def aggb(k):
            call("gatres_gat_agg_bwd", ptr(topo.rowptr), ptr(topo.col), ptr(topo.rowptr_t), ptr(topo.col_t), ptr(g[k]),
                 ptr(h[k]), ptr(ss[k]), ptr(sd[k]), ptr(m[k]), ptr(l[k]), ptr(a_s), ptr(a_d), ptr(rec), ptr(dsd), ptr(dh),
                 ptr(partial), P, S, F * K, F * K + F, F * K + 2 * F, B, N, topo.E1, H, nc, stream())

        def linb(k):
            call("gatres_linear_bwd", ptr(g[k]), ptr(x[k]), ptr(W), None, None, ptr(dx), ptr(partial), P, S, 0, M, K, H,
                 nc, stream())

        proj(0)
        agg(0)                      # real (m, l) so exp() in the backward stays finite
        for k in range(n_sets):
            proj(k)
            agg(k)
        for name, fn, bpn in ((f"linear_att_fwd K={K} H={H}", proj, 4 * K + 4 * F + 8 * H),
                              (f"gat_agg_fwd H={H} C={nc}", agg, 4 * F + 8 * H + 4 * F + 8 * H),
                              (f"gat_agg_bwd (p1+p2) H={H} C={nc}", aggb, 20 * F + 52 * H),
                              (f"linear_bwd (dx+dW) K={K} H={H}", linb, 4 * F + 4 * K + 4 * K)):
            t = time_launches(fn, reps, n_sets)
            rows.append({"kernel": name, "us": t * 1e6, "bytes_per_node": bpn, "GBps": M * bpn / t / 1e9})
        del x, h, out, g, partial
    z, x0, o = rnd(M, nc), rnd(M, nc), rnd(M, nc)

    def mean(k):
        call("gatres_mean_res_fwd", ptr(topo.rowptr), ptr(topo.col), ptr(z[k]), ptr(x0[k]), ptr(o[k]), B, N, nc, stream())

    def meanb(k):
        call("gatres_mean_res_bwd", ptr(topo.rowptr), ptr(topo.rowptr_t), ptr(topo.col_t), ptr(z[k]), None, ptr(o[k]),
             None, B, N, nc, stream())

    for name, fn, bpn in ((f"mean_res_fwd C={nc}", mean, 12 * nc), (f"mean_res_bwd C={nc}", meanb, 8 * nc)):
        t = time_launches(fn, reps, n_sets)
        rows.append({"kernel": name, "us": t * 1e6, "bytes_per_node": bpn, "GBps": M * bpn / t / 1e9})
    return rows


# ----------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a GPU: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
        pg = torch.distributed.group.WORLD

    from gnn_pressure_estimation_b200.GraphModels import GATResMeanConv
    from gnn_pressure_estimation_b200.graph import Topology
    from gnn_pressure_estimation_b200.train_step import TrainStep
    from oracle import gatres_oracle as O      # synthetic-input recipe + cpu_baseline leg only

    nb, nc = model_cfg(args.model)
    ei, N = build_graph(args.graph)
    B, K, W = args.batch, max(1, args.steps), max(3, args.warmup)
    M = B * N
    torch.manual_seed(0)
    model = GATResMeanConv(num_blocks=nb, nc=nc)
    model.load_state_dict(O.make_oracle(nb, nc, seed=0).state_dict())     # identical weights on every rank
    model = model.to(dev)
    topo = Topology.build(ei.to(dev), N)
    mask_count = int(N * MASK_RATE)

    # synthetic batches: a pool of distinct snapshots per rank, pinned on the host and resident on the device
    pool = 8
    host, devp = [], []
    for k in range(pool):
        x, y, mask = O.synthetic_snapshots(N, B, MASK_RATE, seed=1234 + 1000 * rank + k)
        hy, hm = y.reshape(-1).pin_memory(), mask.view(torch.uint8).pin_memory()
        host.append((hy, hm))
        devp.append((hy.to(dev), hm.to(dev)))

    if args.profile_kernels:
        os.environ["GATRES_PROFILE_EAGER"] = "1"
        kernel_table(args.hbm_batch if nc == 32 else max(64, args.hbm_batch // 8), N, topo, nc, True, dev)
        torch.cuda.synchronize()
        return

    if args.mode == "train":
        ts = TrainStep(model, topo, B, mask_count, process_group=pg, use_graph=not args.no_graph)
        ts.capture(warmup=2)

        def step_resident(k):
            y, m = devp[k % pool]
            ts.step(y, y, m)

        loss_host = torch.zeros(K + W, dtype=torch.float32).pin_memory()

        def step_e2e(k, slot):
            y, m = host[k % pool]
            ts.step(y, y, m)                                   # x = y unmasked; the mask is applied on the device
            loss_host[slot:slot + 1].copy_(ts.loss, non_blocking=True)

        launches_per_step = ts.kernels_per_step             # counted by the library while the step was captured
        h2d = M * (4 + 1)           # x (= y, copied once; auxil.py:96-97) + the uint8 mask
        d2h = 4
    else:
        eib = O.collate_edge_index(ei, N, B).to(dev)
        xs = [devp[k][0].view(-1, 1) for k in range(pool)]
        out_host = torch.zeros(M, dtype=torch.float32).pin_memory()
        xdev = torch.empty(M, 1, device=dev)

        def step_resident(k):
            with torch.no_grad():
                model(xs[k % pool], eib)

        def step_e2e(k, slot):
            xdev.copy_(host[k % pool][0].view(-1, 1), non_blocking=True)
            with torch.no_grad():
                out_host.copy_(model(xdev, eib).view(-1), non_blocking=True)

        from gnn_pressure_estimation_b200 import _lib as _gl
        n0 = _gl.load().gatres_launch_count()
        step_resident(0)
        launches_per_step = int(_gl.load().gatres_launch_count() - n0)
        h2d, d2h = 4 * M, 4 * M

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, with_slot):
        for w in range(W):
            fn(w, w) if with_slot else fn(w)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s = torch.cuda.current_stream()
        with ClockSampler(local) as cs:
            e0.record(s)
            for k in range(K):
                fn(k, W + k) if with_slot else fn(k)
            e1.record(s)
            barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms, cs.summary()

    ms_res, clocks = timed(step_resident, False)
    ms_e2e, clocks_e2e = timed(step_e2e, True)
    value = world * B * K / (ms_res * 1e-3)
    e2e_value = world * B * K / (ms_e2e * 1e-3)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_res / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "global_batch": world * B, "parallelism": f"dp{world}",
                       "cuda_graph": not args.no_graph,
                       "kernels": ("snapshot-resident cluster kernels (whole forward / backward stack per launch)"
                                   if args.mode == "train" and launches_per_step < 20 else "layer-by-layer kernels"),
                       "l2": f"no flush: inputs rotate over {pool} resident batches and one step streams "
                             f"~{(ts.saved.numel() + ts.scratch.numel()) * 4 / 1e6:.0f} MB of "
                             "saved activations + scratch (L2 is 126 MB)" if args.mode == "train" else
                             f"no flush: inputs rotate over {pool} resident batches"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / K, "clocks": clocks_e2e},
            "gpu_launches": launches_per_step * K}
    if args.mode == "train":
        line["final_loss"] = float(ts.loss.item())

    if rank == 0 and not args.skip_kernel_leg and args.graph == "ctown":
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak, which = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
        hbm_B = args.hbm_batch if nc == 32 else max(64, args.hbm_batch // 8)
        in_step = kernel_table(B, N, topo, nc, False, dev)
        hbm = kernel_table(hbm_B, N, topo, nc, True, dev)
        dom = max((r for r in hbm if r["kernel"].startswith("gat_agg")), key=lambda r: r["us"])
        dom_l2 = next(r for r in in_step if r["kernel"] == dom["kernel"])
        traffic = None
        try:                                   # DRAM bytes per launch from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_r1.json")))
            if tj.get("hbm_batch") == hbm_B:
                traffic = tj["bytes_per_launch"].get(dom["kernel"])
        except Exception:
            pass
        hbm_regime = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["GBps"], "peak": peak,
                      "unit": "GB/s", "frac": dom["GBps"] / peak, "traffic": traffic, "peak_source": which,
                      "workload": f"{hbm_B} snapshots x {N} nodes per launch (tensors larger than L2), "
                                  f"{dom['bytes_per_node']} algorithmic B/node, {dom['us']:.1f} us/launch"}
        in_step_layer = {"bound": "hbm", "kernel": dom_l2["kernel"], "achieved": dom_l2["GBps"], "peak": peak,
                         "unit": "GB/s", "frac": dom_l2["GBps"] / peak, "traffic": None, "peak_source": which,
                         "workload": f"bench batch ({B} snapshots, L2-resident, {dom_l2['us']:.1f} us/launch)"}
        # `roofline` = the dominant kernel of the TIMED step (the launch list of this command is under profiles/);
        # `roofline_hbm_regime` = the dominant layer kernel where an HBM roofline is meaningful (tensors >> L2).
        line["roofline"] = in_step_layer
        line["roofline_hbm_regime"] = hbm_regime
        if args.mode == "train" and ts.kernels_per_step < 20:
            # the timed step ran the snapshot-resident cluster kernels: its dominant launch is the whole-stack backward
            # (one kernel).  Algorithmic bytes = the per-kernel figures of SURVEY 8d summed over the stack.
            import ctypes as C
            from gnn_pressure_estimation_b200 import _lib as gl
            F1, F2 = 2 * nc, nc
            blk_fwd = (4 * nc + 4 * F1 + 16) + (8 * F1 + 32) + (4 * F1 + 4 * F2 + 8) + (8 * F2 + 16) + 12 * nc
            blk_bwd = (20 * F1 + 104) + (20 * F2 + 52) + (4 * F1 + 8 * nc) + (4 * F2 + 8 * F1) + 8 * nc
            bpn = {"fwd": nb * blk_fwd + 2 * (4 + 4 * nc), "bwd": nb * blk_bwd + (4 + 8 * nc) + (4 + 4 * nc)}
            d = C.byref(ts.desc)
            calls = {
                "fwd": lambda k: gl.call("gatres_forward", d, gl.ptr(ts.flat), gl.ptr(ts.xm), gl.ptr(ts.out), gl.ptr(ts.saved),
                                         gl.ptr(ts.scratch), gl.stream()),
                "bwd": lambda k: gl.call("gatres_backward", d, gl.ptr(ts.flat), gl.ptr(ts.xm), gl.ptr(ts.saved), gl.ptr(ts.d_out),
                                         None, gl.ptr(ts.grads), gl.ptr(ts.scratch), gl.stream()),
            }
            res = {}
            for name in ("fwd", "bwd"):
                t = time_launches(calls[name], 20, 1)
                res[name] = {"us": t * 1e6, "bytes_per_node": bpn[name], "GBps": M * bpn[name] / t / 1e9}
            step_us = ms_res / K * 1e3
            rtraffic = None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_r1.json")))
                if tj.get("resident_batch") == B:
                    rtraffic = tj["bytes_per_launch"].get("resident_bwd_kernel")
            except Exception:
                pass
            line["roofline"] = {
                "bound": "hbm", "kernel": f"resident_bwd_kernel (whole backward stack, {nb} blocks, one launch)",
                "achieved": res["bwd"]["GBps"], "peak": peak, "unit": "GB/s", "frac": res["bwd"]["GBps"] / peak,
                "traffic": rtraffic, "peak_source": which, "share_of_step": res["bwd"]["us"] / step_us,
                "workload": f"bench batch ({B} snapshots): {res['bwd']['us']:.0f} us/launch for "
                            f"{bpn['bwd']} algorithmic B/node (SURVEY 8d per-kernel accounting summed over the stack); "
                            f"forward stack {res['fwd']['us']:.0f} us/launch, {bpn['fwd']} B/node, "
                            f"{res['fwd']['GBps']:.0f} GB/s.  The working set of a layer is L2-resident at this batch: the "
                            "kernel is cluster-barrier / issue bound, not HBM bound (profiles/r1_resident.md); the HBM-regime "
                            "figure of the aggregation kernels is under roofline_hbm_regime"}
        if args.kernels_json:
            os.makedirs(os.path.dirname(os.path.abspath(args.kernels_json)), exist_ok=True)
            json.dump({"in_step_batch": B, "in_step": in_step, "hbm_batch": hbm_B, "hbm": hbm, "peak_gbs": peak},
                      open(args.kernels_json, "w"), indent=1)
    if world > 1:
        torch.distributed.barrier()

    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        cb, _, _ = cpu_reference(args, steps=20, warmup=2, budget_s=25.0)
        line["cpu_baseline"] = cb
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # NCCL communicators referenced by a captured CUDA graph can stall a clean teardown: every rank is
        # done once it passes this barrier, so leave without destroying the group.
        torch.distributed.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
