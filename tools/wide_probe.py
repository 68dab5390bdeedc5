"""Time the wide (nc = 64 / 128) projection kernels on their own and check them against an fp64 product.

    python tools/wide_probe.py [--rows 794624] [--json out.json]

Run once per build / environment (GATRES_TC_WIDE2=0 selects the first-generation K-chunked kernel): CUDA events around
graph-replayed launches on inputs larger than L2, as bench.py does for the aggregation kernels.
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
    ap.add_argument("--json", default=None)
    ap.add_argument("--narrow", action="store_true", help="also time the nc = 32 projections")
    args = ap.parse_args()
    from gnn_pressure_estimation_b200 import _lib, ops  # noqa: F401
    lib = _lib.load()
    dev = torch.device("cuda:0")
    peak = 6540.5
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    out = {"rows": args.rows, "wide2": os.environ.get("GATRES_TC_WIDE2", "1"), "pair": os.environ.get("GATRES_TC_PAIR", "0"), "kernels": []}
    lib.gatres_set_tensor_core(2)
    shapes = [(2, 128, 128), (1, 128, 256), (2, 64, 64), (1, 64, 128)] + ([(2, 32, 32), (1, 32, 64)] if args.narrow else [])
    for H, C, fin in shapes:
        M = args.rows
        g = torch.Generator().manual_seed(5)
        x = torch.randn(M, fin, generator=g).to(dev)
        W = (torch.randn(H * C, fin, generator=g) / fin ** 0.5).to(dev)
        a_s = (torch.randn(H * C, generator=g) * 0.3).to(dev)
        a_d = (torch.randn(H * C, generator=g) * 0.3).to(dev)
        h, ss, sd = torch.ops.gatres.linear_att_fwd(x, W, a_s, a_d, H, C)
        torch.cuda.synchronize()
        # parity on a sample of rows (fp64 product)
        idx = torch.cat([torch.arange(0, 300, device=dev), torch.randint(0, M, (4000,), device=dev),
                         torch.arange(M - 300, M, device=dev)])
        h64 = x[idx].double() @ W.double().T
        s64 = (h64.view(-1, H, C) * a_s.double().view(1, H, C)).sum(-1)
        d64 = (h64.view(-1, H, C) * a_d.double().view(1, H, C)).sum(-1)
        eh = float((h[idx].double() - h64).abs().max() / h64.abs().max())
        es = float((ss.view(M, H)[idx].double() - s64).abs().max() / s64.abs().max())
        ed = float((sd.view(M, H)[idx].double() - d64).abs().max() / d64.abs().max())
        # timing: graph of `iters` launches
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for _ in range(3):
                torch.ops.gatres.linear_att_fwd(x, W, a_s, a_d, H, C)
            st.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=st):
                for _ in range(args.iters):
                    torch.ops.gatres.linear_att_fwd(x, W, a_s, a_d, H, C)
            gr.replay()
            st.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            gr.replay()
            e1.record(st)
            st.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / args.iters
        bytes_row = 4 * (fin + H * C + 2 * H)
        gbps = M * bytes_row / us / 1e3
        rec = {"shape": f"K={fin} N={H * C} H={H}", "us": us, "bytes_per_row": bytes_row, "GBps": gbps, "frac": gbps / peak,
               "tf32x3_TFLOPs": 3 * 2.0 * M * fin * H * C / us / 1e6, "err_h": eh, "err_s_src": es, "err_s_dst": ed}
        print(json.dumps(rec), flush=True)
        out["kernels"].append(rec)
        del x, h, ss, sd
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
