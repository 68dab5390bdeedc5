"""Topology ingestion for the GATRes hot path: EPANET ``.inp`` -> reference-ordered
``edge_index`` without wntr / networkx / PyG, plus the synthetic water-network
generators the benchmarks use (the real ``inputs/ctown.inp`` in the reference
tree is a Git-LFS pointer, /root/reference/inputs/ctown.inp:1-3).

The edge ORDER is part of the drop-in contract ("bit-exact CSR"): the reference
builds its graph template in /root/reference/gnn_pressure_estimation/utils/DataLoader.py:236-258
(``nx.Graph(wn.to_graph()).to_undirected()`` -> ``subgraph(keep).copy()``) and
:28-37 (``from_networkx``).  What those library calls do to adjacency order is
restated here as three explicit steps:

  collapse   MultiDiGraph successors -> simple undirected adjacency, inserting
             ``u-v`` while walking ``for u in nodes: for v in succ[u]``;
  reinsert   every networkx ``copy``/``to_undirected`` rebuilds the adjacency by
             walking ``for u in nodes: for v in adj[u]`` and inserting both
             directions, which can move a later-inserted lower-id neighbour
             forward;
  emit       ``from_networkx``: ``edge_index[:, k] = (id[u], id[v])`` walking
             the same double loop, ids = position in node order.

tests/test_topology.py checks this against real networkx on random multigraphs.
Caveat: when fewer than half of the nodes are kept, networkx's subgraph view walks
the kept-node *set* (string-hash order, different from run to run), so the reference
has no reproducible order to match there; this module always keeps registry order,
which is what networkx does for real networks (C-Town keeps 388 of 396 nodes).
This is host-side, once-per-``.inp`` indexing work (integer only).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

NODE_SECTIONS = ("JUNCTIONS", "RESERVOIRS", "TANKS")      # wntr node registry order
LINK_SECTIONS = ("PIPES", "PUMPS", "VALVES")              # wntr link registry order


@dataclass
class WaterNetwork:
    """Names only — all the hot path needs from an .inp."""
    junctions: List[str]
    reservoirs: List[str]
    tanks: List[str]
    links: List[Tuple[str, str, str]]        # (link id, start node, end node) in registry order
    link_kinds: Optional[List[str]] = None   # section of each link (parallel to ``links``), if known

    @property
    def node_names(self) -> List[str]:
        return self.junctions + self.reservoirs + self.tanks


# --------------------------------------------------------------------------- #
# .inp parsing
# --------------------------------------------------------------------------- #
def parse_inp_text(text: str) -> WaterNetwork:
    """Minimal EPANET .inp reader: ``[SECTION]`` headers (case-insensitive),
    ``;`` starts a comment, first token of node rows is the id, first three of
    link rows are id/start/end.  Link rows are gathered per section and emitted
    in PIPES, PUMPS, VALVES order regardless of their order in the file."""
    if text.startswith("version https://git-lfs"):
        raise ValueError("this .inp is a Git-LFS pointer, not a network file")
    nodes: Dict[str, List[str]] = {s: [] for s in NODE_SECTIONS}
    links: Dict[str, List[Tuple[str, str, str]]] = {s: [] for s in LINK_SECTIONS}
    section = None
    for raw in text.splitlines():
        line = raw.split(";", 1)[0].strip()
        if not line:
            continue
        if line.startswith("["):
            section = line.strip("[]").strip().upper()
            continue
        tok = line.split()
        if section in nodes:
            nodes[section].append(tok[0])
        elif section in links:
            if len(tok) < 3:
                raise ValueError(f"malformed [{section}] row: {raw!r}")
            links[section].append((tok[0], tok[1], tok[2]))
    return WaterNetwork(nodes["JUNCTIONS"], nodes["RESERVOIRS"], nodes["TANKS"],
                        [l for s in LINK_SECTIONS for l in links[s]],
                        [s for s in LINK_SECTIONS for _ in links[s]])


def parse_inp(path: str) -> WaterNetwork:
    with open(path, "r", errors="replace") as f:
        return parse_inp_text(f.read())


def write_inp(wn: WaterNetwork) -> str:
    """Serialise names back to a skeletal .inp (used to exercise the parser)."""
    kinds = list(wn.link_kinds) if wn.link_kinds is not None else ["PIPES"] * len(wn.links)
    out = ["[TITLE]", "synthetic network ; generated", ""]
    for sec, names in zip(NODE_SECTIONS, (wn.junctions, wn.reservoirs, wn.tanks)):
        out.append(f"[{sec}]")
        out.append(";ID  attributes")
        out += [f" {n}\t0\t0\t;" for n in names]
        out.append("")
    for sec in LINK_SECTIONS:
        out.append(f"[{sec}]")
        out += [f" {lid}\t{a}\t{b}\t100\t;" for (lid, a, b), k in zip(wn.links, kinds) if k == sec]
        out.append("")
    out.append("[END]")
    return "\n".join(out)


# --------------------------------------------------------------------------- #
# reference adjacency ordering
# --------------------------------------------------------------------------- #
def _collapse(nodes: Sequence[str], links: Iterable[Tuple[str, str, str]]) -> Dict[str, Dict[str, None]]:
    succ: Dict[str, Dict[str, None]] = {n: {} for n in nodes}
    for _, a, b in links:                       # MultiDiGraph.add_edge(start, end, key=name)
        if a not in succ:
            succ[a] = {}
        if b not in succ:
            succ[b] = {}
        succ[a].setdefault(b)
    adj: Dict[str, Dict[str, None]] = {n: {} for n in succ}
    for u, nb in succ.items():                  # nx.Graph(multidigraph)
        for v in nb:
            adj[u].setdefault(v)
            adj[v].setdefault(u)
    return adj


def _reinsert(adj: Dict[str, Dict[str, None]], keep: Optional[set] = None) -> Dict[str, Dict[str, None]]:
    new: Dict[str, Dict[str, None]] = {n: {} for n in adj if keep is None or n in keep}
    for u, nb in adj.items():
        if u not in new:
            continue
        for v in nb:
            if v in new:
                new[u].setdefault(v)
                new[v].setdefault(u)
    return new


def reference_edge_index_for(wn: WaterNetwork, keep: Optional[Sequence[str]]) -> Tuple[np.ndarray, List[str]]:
    """-> (edge_index int64 [2,E] in the reference's order, kept node names) for an explicit keep list
    (``graph.subgraph(keep_list).copy()``, DataLoader.py:252; ``None`` keeps every node)."""
    adj = _collapse(wn.node_names, wn.links)
    adj = _reinsert(adj)                                        # .to_undirected()
    if keep is not None:
        adj = _reinsert(adj, set(keep))                         # .subgraph(keep).copy(): graph order, filtered
    names = list(adj)
    idx = {n: k for k, n in enumerate(names)}
    src = [idx[u] for u, nb in adj.items() for _ in nb]
    dst = [idx[v] for nb in adj.values() for v in nb]
    return np.asarray([src, dst], dtype=np.int64).reshape(2, -1), names


def reference_edge_index(wn: WaterNetwork, removal: str = "keep_junction") -> Tuple[np.ndarray, List[str]]:
    """-> (edge_index int64 [2,E] in the reference's order, kept node names).

    ``removal`` follows DataLoader.get_keep_list (:40-58): ``keep_junction``
    (default, train.py:598-603) or ``keep_all``.
    """
    if removal == "keep_junction":
        return reference_edge_index_for(wn, wn.junctions)
    if removal == "keep_all":
        return reference_edge_index_for(wn, None)
    raise ValueError(f"unsupported removal {removal!r}")


def edge_index_from_inp(path: str, removal: str = "keep_junction") -> Tuple[np.ndarray, List[str]]:
    return reference_edge_index(parse_inp(path), removal)


# --------------------------------------------------------------------------- #
# synthetic networks
# --------------------------------------------------------------------------- #
def _mst_edges(pts: np.ndarray) -> List[Tuple[int, int]]:
    n = len(pts)
    in_tree = np.zeros(n, bool)
    best = np.full(n, np.inf)
    parent = np.full(n, -1)
    in_tree[0] = True
    d = ((pts - pts[0]) ** 2).sum(1)
    best, parent[:] = d, 0
    best[0] = np.inf
    edges = []
    for _ in range(n - 1):
        k = int(np.argmin(np.where(in_tree, np.inf, best)))
        edges.append((int(parent[k]), k))
        in_tree[k] = True
        d = ((pts - pts[k]) ** 2).sum(1)
        upd = (d < best) & ~in_tree
        best[upd], parent[upd] = d[upd], k
    return edges


def ctown_shaped(n_junctions: int = 388, n_edges: int = 429, seed: int = 0, max_degree: int = 5,
                 with_auxiliaries: bool = True) -> WaterNetwork:
    """A connected planar-ish network with C-Town's published size: 388 junctions
    and 429 distinct junction-junction links (minimum spanning tree over random
    2-D points + shortest extra links under a degree cap).  With
    ``with_auxiliaries`` it also carries 1 reservoir, 7 tanks, parallel pumps and
    valves so ``keep_junction`` and multi-edge collapsing are exercised; the
    kept graph still has exactly ``n_junctions`` nodes / ``n_edges`` edges."""
    rng = np.random.RandomState(seed)
    pts = rng.rand(n_junctions, 2)
    tree = _mst_edges(pts)
    have = {(min(a, b), max(a, b)) for a, b in tree}
    deg = np.zeros(n_junctions, int)
    for a, b in tree:
        deg[a] += 1
        deg[b] += 1
    d2 = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
    iu = np.triu_indices(n_junctions, 1)
    order = np.argsort(d2[iu], kind="stable")
    extra: List[Tuple[int, int]] = []
    for k in order:
        if len(tree) + len(extra) >= n_edges:
            break
        a, b = int(iu[0][k]), int(iu[1][k])
        if (a, b) in have or deg[a] >= max_degree or deg[b] >= max_degree:
            continue
        have.add((a, b))
        extra.append((a, b))
        deg[a] += 1
        deg[b] += 1
    pairs = tree + extra
    assert len(pairs) == n_edges
    perm = rng.permutation(len(pairs))                   # link order in the file is not node order
    flip = rng.rand(len(pairs)) < 0.5
    junctions = [f"J{k + 1}" for k in range(n_junctions)]
    links: List[Tuple[str, str, str]] = []
    kinds: List[str] = []
    n_pump = 11 if with_auxiliaries else 0
    n_valve = 4 if with_auxiliaries else 0
    for r, p in enumerate(perm):
        a, b = pairs[p]
        if flip[p]:
            a, b = b, a
        kind = "PUMPS" if r < n_pump - 2 else ("VALVES" if r < n_pump - 2 + n_valve else "PIPES")
        links.append((f"L{r + 1}", junctions[a], junctions[b]))
        kinds.append(kind)
    reservoirs: List[str] = []
    tanks: List[str] = []
    if with_auxiliaries:
        reservoirs = ["R1"]
        tanks = [f"T{k + 1}" for k in range(7)]
        # two parallel pumps duplicate existing pump links (collapse to one edge, one reversed)
        links.append(("PU_dup1", links[0][1], links[0][2])); kinds.append("PUMPS")
        links.append(("PU_dup2", links[1][2], links[1][1])); kinds.append("PUMPS")
        att = rng.choice(n_junctions, 8, replace=False)
        for k, name in enumerate(reservoirs + tanks):
            links.append((f"PA{k + 1}", name, junctions[int(att[k])])); kinds.append("PIPES")
    wn = WaterNetwork(junctions, reservoirs, tanks, [], [])
    for sec in LINK_SECTIONS:                            # registry order = PIPES, PUMPS, VALVES
        wn.links += [l for l, k in zip(links, kinds) if k == sec]
        wn.link_kinds += [sec] * sum(k == sec for k in kinds)
    return wn


def scaled_wdn(n_nodes: int = 100_000, n_edges: int = 115_000, seed: int = 0) -> WaterNetwork:
    """Scaled synthetic WDN (BASELINE.json config 5): a side x side lattice (+ a
    chain for the remainder), random spanning tree (Kruskal on random weights)
    plus random extra lattice edges up to ``n_edges`` (mean degree 2.3)."""
    rng = np.random.RandomState(seed)
    side = int(np.floor(np.sqrt(n_nodes)))
    n_lat = side * side
    ids = np.arange(n_lat).reshape(side, side)
    cand = np.concatenate([np.stack([ids[:, :-1].ravel(), ids[:, 1:].ravel()], 1),
                           np.stack([ids[:-1, :].ravel(), ids[1:, :].ravel()], 1)])
    cand = cand[rng.permutation(len(cand))]
    parent = np.arange(n_lat)

    def find(a: int) -> int:
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    tree, rest = [], []
    for a, b in cand.tolist():
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[ra] = rb
            tree.append((a, b))
        else:
            rest.append((a, b))
    chain = [(k - 1, k) for k in range(n_lat, n_nodes)]
    need = n_edges - len(tree) - len(chain)
    assert 0 <= need <= len(rest)
    pairs = tree + chain + rest[:need]
    perm = rng.permutation(len(pairs))
    names = [f"N{k}" for k in range(n_nodes)]
    links = [(f"P{r}", names[pairs[p][0]], names[pairs[p][1]]) for r, p in enumerate(perm.tolist())]
    return WaterNetwork(names, [], [], links)


def tiny_network() -> WaterNetwork:
    """7 junctions incl. one isolated node, one tank, a parallel link and a
    reversed duplicate — the smallest case that exercises every ordering rule."""
    j = [f"J{k}" for k in range(1, 8)]
    links = [("P1", "J1", "J2"), ("P2", "J3", "J2"), ("P3", "J3", "J4"), ("P4", "J5", "J4"),
             ("P5", "J6", "J5"), ("P6", "J2", "J5"), ("P7", "T1", "J1"), ("P8", "J6", "J3"),
             ("PU1", "J1", "J2"), ("PU2", "J4", "J3")]
    return WaterNetwork(j, [], ["T1"], links)
