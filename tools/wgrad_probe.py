"""Weight gradient dW = dh^T x of the nc = 32 layers: check against an fp64 product and time (CUDA events around graph replays).

    python tools/wgrad_probe.py [--rows 794624]          # GATRES_WGRAD_TC=0 selects the mma.sync kernel
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2048 * 388)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    from gnn_pressure_estimation_b200 import _lib
    from gnn_pressure_estimation_b200._lib import call, ptr, stream
    lib = _lib.load()
    lib.gatres_set_tensor_core(2)
    dev = torch.device("cuda:0")
    M = args.rows
    for NO, KI in [(64, 32), (32, 64)]:
        g = torch.Generator().manual_seed(3)
        dh = torch.randn(M, NO, generator=g).to(dev)
        x = torch.randn(M, KI, generator=g).to(dev)
        W = torch.zeros(NO, KI, device=dev)
        grads = torch.zeros(NO * KI, device=dev)
        H, C = (2, 32) if NO == 64 else (1, 32)

        def run():
            call("gatres_linear_bwd", ptr(dh), ptr(x), ptr(W), None, None, None, ptr(grads), NO * KI, 0, 0, M, KI, H, C, stream())

        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            run()
            st.synchronize()
            ref = dh.double().T @ x.double()
            err = float((grads.view(NO, KI).double() - ref).abs().max() / ref.abs().max())
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=st):
                for _ in range(args.iters):
                    run()
            gr.replay()
            st.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            gr.replay()
            e1.record(st)
            st.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / args.iters
        bytes_row = 4 * (NO + KI)
        print(json.dumps({"shape": f"dW[{NO}x{KI}]", "wgrad_tc": os.environ.get("GATRES_WGRAD_TC", "1"), "us": us,
                          "GBps": M * bytes_row / us / 1e3, "frac": M * bytes_row / us / 1e3 / 6540.5, "rel_err": err}), flush=True)


def fused(args):
    """conv1 backward of the nc = 32 layers in one call: dx = (dh W + add) * (ref > 0), dW += dh^T x"""
    from gnn_pressure_estimation_b200 import _lib
    from gnn_pressure_estimation_b200._lib import call, ptr, stream
    lib = _lib.load()
    lib.gatres_set_tensor_core(2)
    dev = torch.device("cuda:0")
    for NO, KI in [(64, 32), (32, 64)]:
        fused_shape(args, NO, KI, dev, call, ptr, stream)


def fused_shape(args, NO, KI, dev, call, ptr, stream):
    M = args.rows
    g = torch.Generator().manual_seed(11)
    dh = torch.randn(M, NO, generator=g).to(dev)
    x = torch.randn(M, KI, generator=g).to(dev)
    W = (torch.randn(NO, KI, generator=g) / 8).to(dev)
    add = torch.randn(M, KI, generator=g).to(dev)
    ref = torch.randn(M, KI, generator=g).to(dev)
    dx = torch.empty(M, KI, device=dev)
    grads = torch.zeros(NO * KI, device=dev)

    def run():
        call("gatres_linear_bwd", ptr(dh), ptr(x), ptr(W), ptr(add), ptr(ref), ptr(dx), ptr(grads), NO * KI, 0, 0, M, KI, NO // 32, 32, stream())

    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        run()
        st.synchronize()
        idx = torch.cat([torch.arange(0, 300, device=dev), torch.randint(0, M, (4000,), device=dev), torch.arange(M - 300, M, device=dev)])
        dx64 = (dh[idx].double() @ W.double() + add[idx].double()) * (ref[idx] > 0)
        e_dx = float((dx[idx].double() - dx64).abs().max() / dx64.abs().max())
        w64 = dh.double().T @ x.double()
        e_dw = float((grads.view(NO, KI).double() - w64).abs().max() / w64.abs().max())
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for _ in range(args.iters):
                run()
        gr.replay()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        gr.replay()
        e1.record(st)
        st.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / args.iters
    bytes_row = 4 * (NO + KI + KI + KI + KI)
    print(json.dumps({"kernel": f"linear_bwd dh[.,{NO}] x[.,{KI}] (dx + dW)", "fused2": os.environ.get("GATRES_LINEAR_BWD_FUSED2", "1"), "us": us,
                      "GBps": M * bytes_row / us / 1e3, "frac": M * bytes_row / us / 1e3 / 6540.5, "err_dx": e_dx, "err_dW": e_dw}), flush=True)


if __name__ == "__main__":
    if "--fused" in sys.argv:
        sys.argv.remove("--fused")
        ap = argparse.ArgumentParser()
        ap.add_argument("--rows", type=int, default=2048 * 388)
        ap.add_argument("--iters", type=int, default=20)
        fused(ap.parse_args())
        sys.exit(0)
    main()
