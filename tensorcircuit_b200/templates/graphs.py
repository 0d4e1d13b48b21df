"""Lattice graphs used by the measurement templates (tensorcircuit/templates/graphs.py:15-58):
networkx graphs whose nodes and edges carry a ``weight`` attribute."""

from functools import partial
from typing import Any, Optional, Sequence

import networkx as nx

Graph = Any


def _per_item(v: Any, n: int, default: float) -> Sequence[float]:
    if v is None:
        v = default
    return list(v) if isinstance(v, (list, tuple)) else [v] * n


def Line1D(n: int, node_weight: Optional[Sequence[float]] = None, edge_weight: Optional[Sequence[float]] = None,
           pbc: bool = True) -> Graph:
    """chain of ``n`` sites; with ``pbc`` the closing edge (n-1, 0) takes the weight of the last
    open edge, as in the reference (graphs.py:44-46)"""
    nw, ew = _per_item(node_weight, n, 0.0), _per_item(edge_weight, n, 1.0)
    g = nx.Graph()
    for i in range(n):
        g.add_node(i, weight=nw[i])
    for i in range(n - 1):
        g.add_edge(i, i + 1, weight=ew[i])
    if pbc:
        g.add_edge(n - 1, 0, weight=ew[max(n - 2, 0)])
    return g


def Even1D(n: int, s: int = 0) -> Graph:
    """every second bond of a ring, starting at site ``s``"""
    g = nx.Graph()
    for i in range(n):
        g.add_node(i, weight=1.0)
    for i in range(s, n, 2):
        g.add_edge(i, (i + 1) % n, weight=1.0)
    return g


Odd1D = partial(Even1D, s=1)
