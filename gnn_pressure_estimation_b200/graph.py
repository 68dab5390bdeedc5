"""Shared-topology handling: every snapshot of a batch is the same water network.

PyG's DataLoader collates B copies of one template graph into a block-diagonal
batch (``edge_index = cat(template + b*N)``, /root/reference/gnn_pressure_estimation/train.py:302,
SURVEY.md §A.5).  The kernels want the template once (one CSR in int32) plus B,
so this module recovers (template, N, B) from what the caller passes to
``forward(x, edge_index, batch, edge_attr)`` and caches the device CSR per
template.  Indexing work only; no feature arithmetic happens here.
"""
from __future__ import annotations

from dataclasses import dataclass
from math import gcd
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib

_ops = None


def _get_ops():
    global _ops
    if _ops is None:
        from . import ops  # noqa: F401  (registers torch.ops.gatres)
        _ops = torch.ops.gatres
    return _ops


@dataclass
class Topology:
    """Device-resident CSR pair of one template graph (see gatres_csr_build)."""
    N: int
    E: int                      # directed edges of the template as given
    E1: int                     # edges after the GATConv rewrite (self loops dropped, N appended)
    edge_index: Tensor          # int64 [2,E] template on the device
    rowptr: Tensor
    col: Tensor
    rowptr_t: Tensor
    col_t: Tensor
    dropped_self_loops: int

    @staticmethod
    def build(edge_index: Tensor, num_nodes: int) -> "Topology":
        """One-off per template: builds both CSRs on the device (one host sync to
        read back the counts)."""
        if not edge_index.is_cuda:
            raise _lib.GatresError("Topology.build needs a CUDA edge_index (no CPU path exists)")
        ops = _get_ops()
        ei = edge_index.to(torch.int64).contiguous()
        rowptr, col, rowptr_t, col_t, info = ops.csr_build(ei, int(num_nodes))
        dropped, e1, bad, _ = (int(v) for v in info.tolist())
        if bad:
            raise _lib.GatresError(f"edge_index names {bad} node ids outside [0, {num_nodes})")
        if dropped:
            # GATConv drops them, SimpleConv(mean) would keep them: the two operators would need
            # different structures.  WDN templates come from simple graphs and never have any.
            raise NotImplementedError("template graphs with self-loops are not supported by the fused mean-conv path")
        return Topology(int(num_nodes), ei.size(1), e1, ei, rowptr, col[:e1], rowptr_t, col_t[:e1], dropped)


class TopologyCache:
    """(rows, edge columns) of a collated batch -> (Topology, B).

    First sight of a shape costs one sync (template inference + full check);
    afterwards each call only enqueues a device-side replication check whose
    result poisons the model output with NaN on mismatch — no per-step sync
    (the reference pays one hidden sync per GATConv call).
    """

    def __init__(self) -> None:
        self._by_shape: Dict[Tuple[int, int, int], Tuple[Topology, int]] = {}
        self._templates: Dict[Tuple[int, int], Topology] = {}
        self.mismatch: Optional[Tensor] = None

    def set_template(self, edge_index: Tensor, num_nodes: int) -> Topology:
        topo = Topology.build(edge_index, num_nodes)
        self._templates[(topo.N, topo.E)] = topo
        return topo

    def _infer(self, M: int, edge_index: Tensor, batch: Optional[Tensor]) -> Tuple[int, int]:
        """-> (N, B) such that edge_index is B shifted copies of its first E/B columns."""
        Et = edge_index.size(1)
        if batch is not None:
            N = int((batch == batch[0]).sum())
            cands = [M // N] if N > 0 and M % N == 0 else []
        else:
            g = gcd(M, Et) if Et > 0 else M
            cands = [b for b in range(g, 0, -1) if g % b == 0]
        for B in cands:
            N, E = M // B, Et // B
            if Et % B:
                continue
            if B == 1:
                return N, 1
            t = edge_index[:, :E]
            if int(t.max()) >= N or int(t.min()) < 0:
                continue
            if torch.equal(edge_index.view(2, B, E), (t.unsqueeze(1) + (torch.arange(B, device=t.device) * N).view(1, B, 1))):
                return N, B
        return M, 1

    def resolve(self, x_rows: int, edge_index: Tensor, batch: Optional[Tensor] = None) -> Tuple[Topology, int]:
        key = (x_rows, edge_index.size(1), edge_index.device.index or 0)
        hit = self._by_shape.get(key)
        ops = _get_ops()
        if self.mismatch is None or self.mismatch.device != edge_index.device:
            self.mismatch = torch.zeros(1, dtype=torch.int32, device=edge_index.device)
        if hit is None:
            N, B = self._infer(x_rows, edge_index, batch)
            E = edge_index.size(1) // B
            topo = self._templates.get((N, E))
            if topo is None or not torch.equal(topo.edge_index, edge_index[:, :E]):
                topo = self.set_template(edge_index[:, :E].contiguous(), N)
            hit = (topo, B)
            self._by_shape[key] = hit
        topo, B = hit
        if edge_index.dtype != torch.int64:
            edge_index = edge_index.to(torch.int64)
        ops.check_replicated(edge_index, topo.edge_index, B, topo.N, self.mismatch)
        return hit
