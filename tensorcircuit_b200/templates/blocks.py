"""Circuit building blocks used by VQE/QAOA scripts (tensorcircuit/templates/blocks.py)."""

from __future__ import annotations

from typing import Any, Callable, Optional, Sequence, Tuple

import numpy as np

from ..circuit import Circuit

Tensor = Any


def state_centric(f: Callable[..., Circuit]) -> Callable[..., Tensor]:
    """blocks.py:20-43: function on circuits -> function on state vectors."""

    def wrapper(s: Tensor, *args: Any, **kws: Any) -> Tensor:
        n = int(round(np.log2(int(np.prod(np.shape(s))))))
        c = Circuit(n, inputs=s)
        c = f(c, *args, **kws)
        return c.state()

    return wrapper


def Bell_pair_block(c: Circuit, links: Optional[Sequence[Tuple[int, int]]] = None) -> Circuit:
    """blocks.py:46-68: |00> -> (|01> - |10>)/sqrt(2) on every link."""
    n = c._nqubits
    if links is None:
        links = [(i, i + 1) for i in range(0, n - 1, 2)]
    for a, b in links:
        c.X(a)
        c.H(a)
        c.cnot(a, b)
        c.X(b)
    return c


def example_block(c: Circuit, param: Tensor, nlayers: int = 2, is_split: bool = False) -> Circuit:
    """blocks.py:113-152: H on all, then per layer (ZZ ladder, RX on all); param [2*nlayers, n]."""
    n = c._nqubits
    for i in range(n):
        c.H(i)
    for j in range(nlayers):
        for i in range(n - 1):
            c.exp1(i, i + 1, unitary=np.kron(np.diag([1.0, -1.0]), np.diag([1.0, -1.0])), theta=param[2 * j, i])
        for i in range(n):
            c.rx(i, theta=param[2 * j + 1, i])
    return c


def _sizen(x: Any) -> int:
    """number of entries as user code sees them (a vmap batch axis does not count)"""
    sh = x.shape if hasattr(x, "shape") else np.shape(x)
    return int(np.prod(sh, dtype=np.int64)) if len(sh) else 1


def QAOA_block(c: Circuit, g: Any, paramzz: Tensor, paramx: Tensor, **kws: Any) -> Circuit:
    """blocks.py:84-110: exp(-i theta Z_i Z_j) on graph edges, then rx on nodes.  A single angle is
    shared by all edges (scaled by the edge weight) / all nodes; otherwise edge i takes
    paramzz[i] and node i takes paramx[i], as in the reference."""
    zz = np.kron(np.diag([1.0, -1.0]), np.diag([1.0, -1.0]))
    if _sizen(paramzz) == 1:
        for e1, e2 in g.edges:
            c.exp1(e1, e2, unitary=zz, theta=paramzz * g[e1][e2].get("weight", 1.0), **kws)
    else:
        for i, (e1, e2) in enumerate(g.edges):
            c.exp1(e1, e2, unitary=zz, theta=paramzz[i], **kws)
    if _sizen(paramx) == 1:
        for n in g.nodes:
            c.rx(n, theta=paramx)
    else:
        for i, n in enumerate(g.nodes):
            c.rx(n, theta=paramx[i])
    return c
