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


if __name__ == "__main__":
    main()
