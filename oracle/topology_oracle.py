"""Oracle for the edge-ordering / CSR contract.  TEST INFRASTRUCTURE ONLY.

Runs the reference's actual library call chain with real networkx (installed in
this image; the reference pins 3.1, this image has 3.6.1 — same insertion-order
semantics for the calls involved):

  /root/reference/gnn_pressure_estimation/utils/DataLoader.py:236
      graph = nx.Graph(wn.to_graph(...)).to_undirected()
  DataLoader.py:255   new_graph = graph.subgraph(keep_list).copy()
  DataLoader.py:29    pgu.from_networkx(new_graph)      [ext PyG]

`wn.to_graph()` [ext wntr 1.0.0] and `from_networkx` [ext PyG] are not
installed; their (tiny) bodies are restated below from their published
behaviour: to_graph = MultiDiGraph, all nodes first in registry order, then one
`add_edge(start, end, key=link_name)` per link in registry order; from_networkx
= `G.to_directed()`, node id = position in `G.nodes()`, edge_index filled from
`G.edges()`.  PARITY UNPINNED for those two restated bodies (no reference
fixture exists); the networkx part is the real thing.

CSR oracle: numpy stable sort by target of the GATConv-rewritten edge list
(self-loops dropped, one per node appended last) — SURVEY.md Appendix B step 6.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def reference_pipeline_edge_index(node_names: Sequence[str], junction_names: Sequence[str],
                                  links: Sequence[Tuple[str, str, str]], removal: str = "keep_junction"
                                  ) -> Tuple[np.ndarray, List[str]]:
    import networkx as nx

    G = nx.MultiDiGraph()
    for n in node_names:
        G.add_node(n)
    for lid, a, b in links:
        G.add_edge(a, b, key=lid)
    graph = nx.Graph(G).to_undirected()
    if removal == "keep_junction":
        new_graph = graph.subgraph(list(junction_names)).copy()
    else:
        new_graph = graph
    D = new_graph.to_directed()
    mapping = dict(zip(D.nodes(), range(D.number_of_nodes())))
    ei = np.empty((2, D.number_of_edges()), dtype=np.int64)
    for k, (s, d) in enumerate(D.edges()):
        ei[0, k] = mapping[s]
        ei[1, k] = mapping[d]
    return ei, list(D.nodes())


def rewrite_edges_np(edge_index: np.ndarray, num_nodes: int, add_self_loops: bool = True) -> np.ndarray:
    keep = edge_index[0] != edge_index[1]
    ei = edge_index[:, keep]
    if add_self_loops:
        loops = np.arange(num_nodes, dtype=np.int64)
        ei = np.concatenate([ei, np.stack([loops, loops])], axis=1)
    return ei


def csr_by_target(edge_index: np.ndarray, num_nodes: int, add_self_loops: bool = True
                  ) -> Tuple[np.ndarray, np.ndarray]:
    """-> (rowptr int32 [N+1], col int32 [E']): in-edges of each target in edge-list
    order (stable), i.e. ascending source with the self-loop last for a
    source-sorted list."""
    ei = rewrite_edges_np(edge_index, num_nodes, add_self_loops)
    order = np.argsort(ei[1], kind="stable")
    col = ei[0][order].astype(np.int32)
    rowptr = np.zeros(num_nodes + 1, dtype=np.int32)
    np.cumsum(np.bincount(ei[1], minlength=num_nodes), out=rowptr[1:])
    return rowptr, col


def csr_by_source(edge_index: np.ndarray, num_nodes: int, add_self_loops: bool = True
                  ) -> Tuple[np.ndarray, np.ndarray]:
    """Transposed structure: out-edges of each source (targets), stable."""
    ei = rewrite_edges_np(edge_index, num_nodes, add_self_loops)
    order = np.argsort(ei[0], kind="stable")
    col = ei[1][order].astype(np.int32)
    rowptr = np.zeros(num_nodes + 1, dtype=np.int32)
    np.cumsum(np.bincount(ei[0], minlength=num_nodes), out=rowptr[1:])
    return rowptr, col
