"""Generates the committed golden fixtures from the CPU oracle (run from the repo
root: `python tests/golden/make_golden.py`).  PARITY UNPINNED by the reference
(it ships no vectors and PyG is not installable here): these pin the ORACLE, so
that the CUDA path, the oracle and future edits of either are all compared with
one frozen set of numbers.  Edge-order fixtures come from real networkx
(oracle/topology_oracle.py)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gnn_pressure_estimation_b200 import topology as T  # noqa: E402
from oracle import gatres_oracle as O  # noqa: E402
from oracle import topology_oracle as TO  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def model_case(name, wn, num_blocks, nc, B, full_grads):
    ei_np, names = TO.reference_pipeline_edge_index(wn.node_names, wn.junctions, wn.links)
    N = len(names)
    ei = torch.from_numpy(ei_np)
    model = O.make_oracle(num_blocks, nc, seed=0)
    x, y, mask = O.synthetic_snapshots(N, B, mask_rate=0.95 if N > 20 else 0.5, seed=1234)
    eib = O.collate_edge_index(ei, N, B)
    out, loss, grads = O.train_step_loss_and_grads(model, x, y, mask, eib)
    case = dict(name=name, num_blocks=num_blocks, nc=nc, B=B, N=N, seed=0, edge_index=ei, x=x, y=y, mask=mask,
                out=out, loss=loss,
                grad_norms={k: float(g.norm()) for k, g in grads.items()},
                grad_heads={k: g.reshape(-1)[:8].clone() for k, g in grads.items()})
    if full_grads:
        case["grads"] = grads
        case["state_dict"] = {k: v.clone() for k, v in model.state_dict().items()}
    torch.save(case, os.path.join(HERE, f"{name}.pt"))
    print(name, "N", N, "E", ei.size(1), "loss", float(loss), "bytes", os.path.getsize(os.path.join(HERE, f"{name}.pt")))


def topo_case(name, wn):
    ei, names = TO.reference_pipeline_edge_index(wn.node_names, wn.junctions, wn.links)
    rp, col = TO.csr_by_target(ei, len(names))
    rpt, colt = TO.csr_by_source(ei, len(names))
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), edge_index=ei, rowptr=rp, col=col, rowptr_t=rpt, col_t=colt,
                        n=len(names))
    print(name, ei.shape)


if __name__ == "__main__":
    torch.set_num_threads(4)
    topo_case("topo_tiny", T.tiny_network())
    topo_case("topo_ctown_shaped", T.ctown_shaped())
    model_case("tiny_2b_32c_B3", T.tiny_network(), 2, 32, 3, True)
    model_case("ctown_small_15b_32c_B8", T.ctown_shaped(), 15, 32, 8, True)
    model_case("ctown_mid_3b_64c_B2", T.ctown_shaped(), 3, 64, 2, False)
    model_case("ctown_large_25b_128c_B2", T.ctown_shaped(), 25, 128, 2, False)
