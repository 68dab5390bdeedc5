#!/usr/bin/env python
"""Benchmark of the GATRes hot path (contract: see DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
  python bench.py --impl reference [--gpus N] [--steps K] ...    # CPU reference arm (oracle port)

Headline workload = BASELINE.json configs[1]: gatres_small training on the C-Town-shaped graph, batch 32 snapshots
per GPU (weak scaling over --gpus), synthetic snapshots, mask_rate 0.95, Adam.  A step = mask + forward + masked MSE +
backward (+ gradient all-reduce) + Adam.  One JSON line on stdout (rank 0).

Timing: W >= 3 warm-up steps, then R >= 5 repeats of EXACTLY K steps; every repeat is bracketed by a barrier +
synchronize on both sides, starts with a GPU-side rendezvous (an all-reduce enqueued right before the first event, so
no rank's clock starts before every rank's stream has arrived), is timed with CUDA events on the launch stream and
reduced with MAX over ranks.  `ms_per_step` / `value` come from the MEDIAN repeat (all repeats are listed under
`timing`); the NVML clock sampler is built before the first barrier and runs across all repeats.

The other BASELINE.json configurations (configs[2..4]) run as short legs after the headline, on every N, and are
attached under `configs`.
"""
from __future__ import annotations

import argparse
import json
import math
import os
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
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--repeats", type=int, default=0, help="timed repeats of the K steps (0 = auto: >= 5, ~0.4 s in total)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="snapshots per GPU per step (configs[1]: 32)")
    ap.add_argument("--model", default="gatres_small", choices=["gatres_small", "gatres_large"])
    ap.add_argument("--graph", default="ctown", choices=["ctown", "scaled"])
    ap.add_argument("--mode", default="train", choices=["train", "infer"])
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-kernel-leg", action="store_true")
    ap.add_argument("--skip-config-legs", action="store_true", help="do not run BASELINE configs[2..4] after the headline")
    ap.add_argument("--profile-kernels", action="store_true",
                    help="only launch each hot kernel at the HBM-regime batch (for ncu); prints nothing")
    ap.add_argument("--hbm-batch", type=int, default=2048, help="batch of the HBM-regime kernel measurement")
    ap.add_argument("--kernels-json", default=None, help="write the per-kernel table here")
    return ap.parse_args()


def model_cfg(name):
    return (15, 32) if name == "gatres_small" else (25, 128)


_GRAPHS = {}


def build_graph(kind):
    if kind not in _GRAPHS:
        from gnn_pressure_estimation_b200 import topology as T
        wn = T.ctown_shaped() if kind == "ctown" else T.scaled_wdn()
        ei, names = T.reference_edge_index(wn)
        _GRAPHS[kind] = (torch.from_numpy(ei), len(names))
    return _GRAPHS[kind]


def workload_name(model, graph, mode, batch):
    g = "C-Town-shaped synthetic graph (N=388, E=858 directed)" if graph == "ctown" else \
        "scaled synthetic WDN (N=100000, E=230000 directed)"
    what = "training step (mask+fwd+MSE+bwd+Adam)" if mode == "train" else "inference forward"
    return f"{model} {what}, {g}, batch {batch} snapshots/GPU, mask_rate {MASK_RATE}"


def config_dict(args, world, B, extra=None):
    """same keys on both arms (the driver compares them)"""
    c = {"workload": workload_name(args.model, args.graph, args.mode, B), "global_batch": world * B,
         "parallelism": f"dp{world}", "cuda_graph": None, "kernels": None, "l2": None}
    c.update(extra or {})
    return c


# ----------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock + throttle reasons while the timed region runs (NVML).  Constructing it initialises NVML
    (serialised across processes): build it BEFORE the barrier that opens a timed region."""
    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown"}
    NOTE = {0x4: "sw_power_cap"}

    def __init__(self, index):
        self.samples, self.reasons, self._stop = [], set(), threading.Event()
        self.max_mhz = None
        self._on = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self._sample()                                   # first call pays the lazy initialisation
            self.samples.clear()
            self.reasons.clear()
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)
        if self.nv is not None:
            self.t.start()

    def _sample(self):
        self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
        r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for bit, name in {**self.BAD, **self.NOTE}.items():
            if r & bit:
                self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            if self._on:
                try:
                    self._sample()
                except Exception:
                    pass
            time.sleep(0.002)

    def __enter__(self):
        self._on = True
        return self

    def __exit__(self, *a):
        self._on = False

    def close(self):
        self._stop.set()
        if self.nv is not None:
            self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}

    def reset(self):
        self.samples, self.reasons = [], set()


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
            float(loss.detach())
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
    world = max(1, args.gpus)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # the CPU arm always runs ONE rank's batch on the host cores (rank 0 only), whatever --gpus says
            "config": config_dict(args, world, args.batch, {"cuda_graph": False, "kernels": "CPU oracle port of the PyG op sequence "
                                                            f"(batch {B} per step on the host cores)", "l2": "n/a (host)"}),
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
    """GB/s of each hot kernel of one block on [B*N] rows.  `bytes_per_node` is the kernel's OWN algorithmic traffic
    (every tensor it must consume / produce counted once): the fused snapshot-tile backward moves
    12 S + 28 H bytes per node (h, g read once, dh written; csrc/gat_agg_tile.cu), the two gather passes 20 S + 52 H;
    `bytes_per_node_unfused` always carries the two-pass figure as the labelled secondary."""
    from gnn_pressure_estimation_b200 import _lib as gl
    from gnn_pressure_estimation_b200 import ops as gops
    from gnn_pressure_estimation_b200._lib import call, ptr, stream
    M = B * N
    f = dict(dtype=torch.float32, device=dev)
    n_sets = 2 if hbm_regime else 4
    reps = 20 if hbm_regime else 60
    rows = []
    tile = nc == 32 and B >= int(gl.load().gatres_set_tile_min_batch(-1))      # snapshot-tile kernels (slab fits for C-Town, nc = 32)

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
        two_pass = 20 * F + 52 * H
        for name, fn, bpn, unf in ((f"linear_att_fwd K={K} H={H}", proj, 4 * K + 4 * F + 8 * H, None),
                                   (f"gat_agg_fwd H={H} C={nc}", agg, 4 * F + 8 * H + 4 * F + 8 * H, None),
                                   (f"gat_agg_bwd H={H} C={nc}", aggb, (12 * F + 28 * H) if tile else two_pass, two_pass),
                                   (f"linear_bwd (dx+dW) K={K} H={H}", linb, 4 * F + 4 * K + 4 * K, None)):
            t = time_launches(fn, reps, n_sets)
            row = {"kernel": name, "us": t * 1e6, "bytes_per_node": bpn, "GBps": M * bpn / t / 1e9}
            if unf is not None:
                row["accounting"] = "fused snapshot-tile kernel (12S+28H)" if tile else "two gather passes (20S+52H)"
                row["bytes_per_node_unfused"] = unf
                row["GBps_unfused_accounting"] = M * unf / t / 1e9
            rows.append(row)
        del x, h, out, g, partial
    z, x0, o = rnd(M, nc), rnd(M, nc), rnd(M, nc)

    def mean(k):
        call("gatres_mean_res_fwd", ptr(topo.rowptr), ptr(topo.col), ptr(z[k]), ptr(x0[k]), ptr(o[k]), B, N, nc, stream())

    def meanb(k):
        call("gatres_mean_res_bwd_e1", ptr(topo.rowptr), ptr(topo.rowptr_t), ptr(topo.col_t), topo.E1, ptr(z[k]), ptr(o[k]),
             B, N, nc, stream())

    for name, fn, bpn in ((f"mean_res_fwd C={nc}", mean, 12 * nc), (f"mean_res_bwd C={nc}", meanb, 8 * nc)):
        t = time_launches(fn, reps, n_sets)
        rows.append({"kernel": name, "us": t * 1e6, "bytes_per_node": bpn, "GBps": M * bpn / t / 1e9})
    return rows


# ----------------------------------------------------------------------------
# one workload = (model, graph, mode, snapshots per GPU)
# ----------------------------------------------------------------------------
class Workload:
    def __init__(self, model_name, graph, mode, B, dev, rank, world, pg, use_graph=True, pool=8, chunks=1):
        from gnn_pressure_estimation_b200.GraphModels import GATResMeanConv
        from gnn_pressure_estimation_b200.graph import Topology
        from gnn_pressure_estimation_b200.train_step import TrainStep
        from gnn_pressure_estimation_b200 import _lib as gl
        from oracle import gatres_oracle as O      # synthetic-input recipe (+ the cpu_baseline leg) only

        self.model_name, self.graph, self.mode, self.B, self.world = model_name, graph, mode, B, world
        self.chunks = chunks                       # inference: forward calls per step (a rank's share in bounded batches)
        nb, nc = model_cfg(model_name)
        ei, N = build_graph(graph)
        self.N, self.nb, self.nc = N, nb, nc
        M = self.M = B * N
        torch.manual_seed(0)
        model = GATResMeanConv(num_blocks=nb, nc=nc)
        model.load_state_dict(O.make_oracle(nb, nc, seed=0).state_dict())     # identical weights on every rank
        self.model = model = model.to(dev)
        self.topo = topo = Topology.build(ei.to(dev), N)
        # synthetic batches: a pool of distinct snapshots per rank, pinned on the host and resident on the device
        pool = max(2, min(pool, (1 << 28) // max(1, M)))                      # bound host memory for the huge legs
        self.pool = pool
        gen = np.random.RandomState(1234 + 1000 * rank)
        self.host, self.devp = [], []
        cnt = int(N * MASK_RATE)
        for k in range(pool):
            if M <= 1 << 22:
                _, y, mask = O.synthetic_snapshots(N, B, MASK_RATE, seed=1234 + 1000 * rank + k)
                y, mask = y.reshape(-1), mask.view(torch.uint8).reshape(-1)
            else:                                                             # same recipe, vectorised for big batches
                y = torch.from_numpy(gen.standard_normal(M).astype(np.float32))
                keys = torch.from_numpy(gen.random_sample((B, N)).astype(np.float32))
                mask = torch.zeros(B, N, dtype=torch.uint8)
                mask.scatter_(1, keys.topk(cnt, dim=1).indices, 1)
                mask = mask.reshape(-1)
            hy, hm = y.pin_memory(), mask.pin_memory()
            self.host.append((hy, hm))
            self.devp.append((hy.to(dev), hm.to(dev)))
        if mode == "train":
            ts = self.ts = TrainStep(model, topo, B, cnt, process_group=pg, use_graph=use_graph)
            if world > 1:
                torch.distributed.barrier()        # ranks build their inputs at different speeds: enter the first
            ts.capture(warmup=2)                   # peer-synchronised step together
            self.loss_host = torch.zeros(4096, dtype=torch.float32).pin_memory()
            self.launches_per_step = ts.kernels_per_step             # counted by the library while the step was captured
            self.h2d, self.d2h = M * (4 + 1), 4     # x (= y, copied once; auxil.py:96-97) + the uint8 mask; the loss
            self.stream_mb = (ts.saved.numel() + ts.scratch.numel()) * 4 / 1e6
        else:
            self.ts = None
            self.eib = O.collate_edge_index(ei, N, B).to(dev)
            self.xs = [self.devp[k][0].view(-1, 1) for k in range(pool)]
            self.out_host = torch.zeros(M, dtype=torch.float32).pin_memory()
            self.xdev = torch.empty(M, 1, device=dev)
            n0 = gl.load().gatres_launch_count()
            self.step_resident(0)
            self.launches_per_step = int(gl.load().gatres_launch_count() - n0)
            self.h2d, self.d2h = 4 * M * chunks, 4 * M * chunks
            self.stream_mb = None

    def step_resident(self, k):
        if self.ts is not None:
            y, m = self.devp[k % self.pool]
            self.ts.step(y, y, m)
        else:
            with torch.no_grad():
                for c in range(self.chunks):
                    self.model(self.xs[(k * self.chunks + c) % self.pool], self.eib)

    def step_e2e(self, k):
        if self.ts is not None:
            y, m = self.host[k % self.pool]
            self.ts.step(y, y, m)                                   # x = y unmasked; the mask is applied on the device
            self.loss_host[k % 4096:k % 4096 + 1].copy_(self.ts.loss, non_blocking=True)
        else:
            for c in range(self.chunks):
                self.xdev.copy_(self.host[(k * self.chunks + c) % self.pool][0].view(-1, 1), non_blocking=True)
                with torch.no_grad():
                    self.out_host.copy_(self.model(self.xdev, self.eib).view(-1), non_blocking=True)

    def kernels(self):
        if self.ts is not None and self.launches_per_step < 20:
            return "snapshot-resident cluster kernels (whole forward / backward stack per launch)"
        if self.ts is None and self.launches_per_step < 4:
            return "snapshot-resident cluster kernel (whole forward stack in one launch)"
        return "layer-by-layer kernels"


class Timer:
    """R repeats of exactly K steps (module docstring)."""

    def __init__(self, dev, local, world):
        self.dev, self.world = dev, world
        self.clock = ClockSampler(local)                      # NVML init happens here, outside any timed region
        self._rv = torch.zeros(1, device=dev)

    def barrier(self):
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def _one(self, fn, K):
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s = torch.cuda.current_stream()
        if self.world > 1:
            torch.distributed.all_reduce(self._rv)            # GPU-side rendezvous: the stream waits for every rank
        e0.record(s)
        for k in range(K):
            fn(k)
        e1.record(s)
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def run(self, fn, K, W, repeats=0, target_ms=400.0, max_repeats=40):
        for w in range(W):
            fn(w)
        est = self._one(fn, K)                                # untimed extra repeat: sizes the repeat count
        if repeats <= 0:
            repeats = int(min(max_repeats, max(5, math.ceil(target_ms / max(est, 1e-3)))))
        if self.world > 1:                                    # every rank must run the same number of repeats
            t = torch.tensor([repeats], device=self.dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            repeats = int(t.item())
        self.clock.reset()
        with self.clock:
            ms = [self._one(fn, K) for _ in range(repeats)]
        med = float(np.median(ms))
        return med, {"repeats": repeats, "repeat_ms": [round(v, 4) for v in ms], "statistic": "median of repeats, max over ranks per repeat"}, self.clock.summary()


def run_leg(timer, name, model_name, graph, mode, B, steps, warmup, dev, rank, world, pg, scaling, note, chunks=1):
    """one short leg of another BASELINE.json configuration (value resident, e2e through the public API)"""
    t0 = time.time()
    try:
        wl = Workload(model_name, graph, mode, B, dev, rank, world, pg, pool=4, chunks=chunks)
        ms, timing, clocks = timer.run(wl.step_resident, steps, warmup, repeats=5)
        ms_e, _, _ = timer.run(wl.step_e2e, steps, warmup, repeats=3)
        out = {"config": name, "workload": workload_name(model_name, graph, mode, B), "note": note,
               "global_batch": world * B * chunks, "n_gpus": world, "scaling": scaling,
               "value": world * B * chunks * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
               "e2e": {"value": world * B * chunks * steps / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": wl.h2d,
                       "d2h_bytes_per_step": wl.d2h},
               "gpu_launches_per_step": wl.launches_per_step, "kernels": wl.kernels(), "timing": timing, "clocks": clocks,
               "wall_s": None}
        del wl
    except Exception as e:                                     # a leg must never take the headline line down
        out = {"config": name, "error": repr(e)[:300]}
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    out["wall_s"] = round(time.time() - t0, 1)
    return out


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

    nb, nc = model_cfg(args.model)
    B, K, W = args.batch, max(1, args.steps), max(3, args.warmup)
    timer = Timer(dev, local, world)

    if args.profile_kernels:
        from gnn_pressure_estimation_b200.graph import Topology
        ei, N = build_graph(args.graph)
        topo = Topology.build(ei.to(dev), N)
        os.environ["GATRES_PROFILE_EAGER"] = "1"
        kernel_table(args.hbm_batch if nc == 32 else max(64, args.hbm_batch // 8), N, topo, nc, True, dev)
        torch.cuda.synchronize()
        return

    wl = Workload(args.model, args.graph, args.mode, B, dev, rank, world, pg, use_graph=not args.no_graph)
    N, M, topo, ts = wl.N, wl.M, wl.topo, wl.ts
    ms_res, timing, clocks = timer.run(wl.step_resident, K, W, args.repeats)
    ms_e2e, timing_e2e, clocks_e2e = timer.run(wl.step_e2e, K, W, args.repeats)
    value = world * B * K / (ms_res * 1e-3)
    e2e_value = world * B * K / (ms_e2e * 1e-3)

    l2 = (f"no flush: inputs rotate over {wl.pool} resident batches and one step streams ~{wl.stream_mb:.0f} MB of "
          "saved activations + scratch (L2 is 126 MB)") if args.mode == "train" else \
        f"no flush: inputs rotate over {wl.pool} resident batches"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_res / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, world, B, {"cuda_graph": not args.no_graph, "kernels": wl.kernels(), "l2": l2}),
            "timing": timing, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": wl.h2d, "d2h_bytes_per_step": wl.d2h,
                    "ms_per_step": ms_e2e / K, "clocks": clocks_e2e, "timing": timing_e2e},
            "gpu_launches": wl.launches_per_step * K}
    if args.mode == "train":
        line["final_loss"] = float(ts.loss.item())

    if rank == 0 and not args.skip_kernel_leg and args.graph == "ctown":
        roofline_leg(args, line, wl, ms_res / K, dev)
    del wl, ts
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    if world > 1:
        torch.distributed.barrier()

    # ---- BASELINE.json configs[2..4] as short legs, on every N (rank 0 reports) -------------------------------------
    headline = args.model == "gatres_small" and args.graph == "ctown" and args.mode == "train" and args.batch == 32
    if headline and not args.skip_config_legs:
        legs = []
        if 1024 % world == 0:
            legs.append(run_leg(timer, "configs[2]", "gatres_small", "ctown", "train", 1024 // world, 10, 3, dev, rank, world, pg,
                                "strong", "global batch 1024 data-parallel (1024 / N snapshots per GPU), gradient all-reduce fused "
                                          "into Adam over NVLink peer memory"))
        if 16384 % world == 0:
            share = 16384 // world
            b3 = min(2048, share)
            legs.append(run_leg(timer, "configs[3]", "gatres_large", "ctown", "infer", b3, 2, 3, dev, rank, world, pg,
                                "strong", "largest GATRes (25 blocks x 128 channels) inference, 16384 snapshots sharded over the "
                                          f"GPUs ({share} per GPU, forward calls of {b3}), no communication", chunks=share // b3))
        legs.append(run_leg(timer, "configs[4] train", "gatres_small", "scaled", "train", 4, 5, 3, dev, rank, world, pg, "weak",
                            "scaled synthetic WDN (100 000 junctions), 4 snapshots per GPU"))
        legs.append(run_leg(timer, "configs[4] infer", "gatres_small", "scaled", "infer", 16, 5, 3, dev, rank, world, pg, "weak",
                            "scaled synthetic WDN (100 000 junctions), 16 snapshots per GPU, no communication"))
        line["configs"] = legs

    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        cb, _, _ = cpu_reference(args, steps=20, warmup=2, budget_s=25.0)
        line["cpu_baseline"] = cb
    timer.clock.close()
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


def load_traffic():
    """DRAM bytes per launch from the committed ncu --set full captures (newest round first)"""
    for name in ("traffic_r2.json", "traffic_r1.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name))), name
        except Exception:
            continue
    return {}, None


def roofline_leg(args, line, wl, step_ms, dev):
    import ctypes as C
    from gnn_pressure_estimation_b200 import _lib as gl
    nb, nc = model_cfg(args.model)
    B, N, M, topo, ts = wl.B, wl.N, wl.M, wl.topo, wl.ts
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, which = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    hbm_B = args.hbm_batch if nc == 32 else max(64, args.hbm_batch // 8)
    in_step = kernel_table(B, N, topo, nc, False, dev)
    hbm = kernel_table(hbm_B, N, topo, nc, True, dev)
    tj, tj_name = load_traffic()
    per_launch = tj.get("bytes_per_launch", {}) if tj.get("hbm_batch") == hbm_B else {}

    def entry(r, workload):
        tr = per_launch.get(r["kernel"])
        alg = M_of[id(r)] * r["bytes_per_node"]
        e = {"bound": "hbm", "kernel": r["kernel"], "achieved": r["GBps"], "peak": peak, "unit": "GB/s",
             "frac": r["GBps"] / peak, "traffic": tr, "traffic_over_algorithmic": (tr / alg) if tr else None,
             "peak_source": which, "accounting": r.get("accounting", "per-kernel algorithmic bytes (SURVEY 8d)"),
             "workload": workload + f", {r['bytes_per_node']} algorithmic B/node, {r['us']:.1f} us/launch"}
        if "GBps_unfused_accounting" in r:
            e["secondary_unfused_accounting"] = {"bytes_per_node": r["bytes_per_node_unfused"],
                                                 "achieved": r["GBps_unfused_accounting"],
                                                 "frac": r["GBps_unfused_accounting"] / peak,
                                                 "label": "two-pass figure of SURVEY 8d (20S+52H) over the fused kernel's time: NOT its traffic"}
        return e

    M_of = {id(r): hbm_B * N for r in hbm}
    M_of.update({id(r): B * N for r in in_step})
    aggs = [r for r in hbm if r["kernel"].startswith("gat_agg")]
    dom = max(aggs, key=lambda r: r["us"])
    dom_l2 = next(r for r in in_step if r["kernel"] == dom["kernel"])
    wl_hbm = f"{hbm_B} snapshots x {N} nodes per launch (tensors larger than L2)"
    # `roofline` = the dominant kernel of the TIMED step; `roofline_hbm_regime` = the aggregation kernels where an
    # HBM roofline is meaningful (tensors >> L2): the slowest one first, every aggregation kernel listed.
    line["roofline"] = entry(dom_l2, f"bench batch ({B} snapshots, L2-resident)")
    line["roofline_hbm_regime"] = entry(dom, wl_hbm)
    line["roofline_hbm_regime"]["aggregation_kernels"] = [
        {"kernel": r["kernel"], "us": r["us"], "bytes_per_node": r["bytes_per_node"], "GBps": r["GBps"], "frac": r["GBps"] / peak,
         "accounting": r.get("accounting", "per-kernel algorithmic bytes")} for r in aggs]
    # the dense contractions of the same layers (tcgen05 3xTF32 projections and their backward) and the mean kernels, same regime
    line["roofline_hbm_regime"]["other_kernels"] = [
        {"kernel": r["kernel"], "us": r["us"], "bytes_per_node": r["bytes_per_node"], "GBps": r["GBps"], "frac": r["GBps"] / peak,
         "accounting": "per-kernel algorithmic bytes"} for r in hbm if not r["kernel"].startswith("gat_agg")]
    if args.mode == "train" and ts is not None and ts.kernels_per_step < 20:
        # the timed step ran the snapshot-resident cluster kernels: its dominant launch is the whole-stack backward
        # (one kernel).  FUSED accounting: what that one launch must move through HBM — the saved activations of every
        # block read once (6 nc + 12 floats per node: h1 ss1 sd1 m1 l1 y1 h2 ss2 sd2 m2 l2 xout), the encoder output,
        # x and d_out, the parameters once and the gradient buffer once.  Everything else (running gradient, dz, dy1,
        # per-row records) is exchanged inside the cluster and is not algorithmic traffic.
        P = ts.P
        per_node_bwd = nb * (6 * nc + 12) * 4 + 4 * nc + 8
        per_node_fwd = nb * (6 * nc + 12) * 4 + 4 * nc + 8                    # the same tensors, written once
        bytes_bwd = M * per_node_bwd + 8 * P
        bytes_fwd = M * per_node_fwd + 4 * P
        F1, F2 = 2 * nc, nc
        blk_bwd_unf = (20 * F1 + 104) + (20 * F2 + 52) + (4 * F1 + 8 * nc) + (4 * F2 + 8 * F1) + 8 * nc
        unf_bwd = M * (nb * blk_bwd_unf + (4 + 8 * nc) + (4 + 4 * nc))
        d = C.byref(ts.desc)
        calls = {
            "fwd": lambda k: gl.call("gatres_forward", d, gl.ptr(ts.flat), gl.ptr(ts.xm), gl.ptr(ts.out), gl.ptr(ts.saved),
                                     gl.ptr(ts.scratch), gl.stream()),
            "bwd": lambda k: gl.call("gatres_backward", d, gl.ptr(ts.flat), gl.ptr(ts.xm), gl.ptr(ts.saved), gl.ptr(ts.d_out),
                                     None, gl.ptr(ts.grads), gl.ptr(ts.scratch), gl.stream()),
        }
        t_f = time_launches(calls["fwd"], 20, 1)
        t_b = time_launches(calls["bwd"], 20, 1)     # includes the 263 KB memset of the gradient buffer
        rtraffic = tj.get("bytes_per_launch", {}).get("resident_bwd_kernel") if tj.get("resident_batch") == B else None
        ach = bytes_bwd / t_b / 1e9
        line["roofline"] = {
            "bound": "hbm", "kernel": f"res2::bwd_kernel (whole backward stack, {nb} blocks, one launch)",
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": rtraffic, "traffic_over_algorithmic": (rtraffic / bytes_bwd) if rtraffic else None,
            "traffic_source": tj_name, "peak_source": which, "share_of_step": t_b * 1e3 / step_ms,
            "accounting": "fused: saved activations read once + inputs + parameters + gradients",
            "algorithmic_bytes_per_launch": bytes_bwd,
            "regime": "latency-bound, L2-resident: one layer's working set is ~3 MB at this batch, the launch is bound by its "
                      "dependent phases (cluster exchange + barriers), not by HBM; the HBM-regime figures of the aggregation "
                      "kernels are under roofline_hbm_regime",
            "workload": f"bench batch ({B} snapshots): {t_b * 1e6:.0f} us/launch for {per_node_bwd} algorithmic B/node",
            "forward_stack": {"kernel": "res2::fwd_tc_kernel<train> (tcgen05 projections, saved activations by bulk copies)", "us": t_f * 1e6, "algorithmic_bytes_per_launch": bytes_fwd,
                              "achieved": bytes_fwd / t_f / 1e9, "frac": bytes_fwd / t_f / 1e9 / peak},
            "secondary_unfused_accounting": {"bytes_per_launch": unf_bwd, "achieved": unf_bwd / t_b / 1e9,
                                             "frac": unf_bwd / t_b / 1e9 / peak,
                                             "label": "sum of the per-kernel figures of SURVEY 8d for the unfused layer kernels: "
                                                      "NOT this kernel's traffic (round 1 reported this one)"}}
    if args.kernels_json:
        os.makedirs(os.path.dirname(os.path.abspath(args.kernels_json)), exist_ok=True)
        json.dump({"in_step_batch": B, "in_step": in_step, "hbm_batch": hbm_B, "hbm": hbm, "peak_gbs": peak},
                  open(args.kernels_json, "w"), indent=1)


if __name__ == "__main__":
    main()
