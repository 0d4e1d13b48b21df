"""Minimal Node / Edge graph with pairwise contraction (see package docstring)."""

import numpy as np


class Edge:
    def __init__(self, node1, axis1, node2=None, axis2=None, name=None):
        self.node1, self.axis1, self.node2, self.axis2 = node1, axis1, node2, axis2
        self.name = name or "__unnamed_edge__"
        self.is_disabled = False

    def is_dangling(self):
        return self.node2 is None

    def is_trace(self):
        return self.node2 is not None and self.node1 is self.node2

    @property
    def dimension(self):
        return self.node1.tensor.shape[self.axis1]

    def get_nodes(self):
        return [self.node1, self.node2]

    def __xor__(self, other):
        return connect(self, other)

    def __repr__(self):
        return "Edge(Dangling Edge)[%d]" % self.axis1 if self.is_dangling() else "Edge(%s[%d] -> %s[%d])" % (self.node1.name, self.axis1, self.node2.name, self.axis2)


class AbstractNode:
    pass


class Node(AbstractNode):
    def __init__(self, tensor, name=None, axis_names=None, backend=None):
        if isinstance(tensor, Node):
            tensor = tensor.tensor
        self.tensor = tensor
        self.name = name or "__unnamed_node__"
        self.edges = [Edge(self, i) for i in range(len(np.shape(tensor)))]
        self.backend = backend

    # -- structure --------------------------------------------------------------------
    @property
    def tensor(self):
        return self._tensor

    @tensor.setter
    def tensor(self, t):
        self._tensor = t

    @property
    def shape(self):
        return tuple(np.shape(self.tensor))

    def get_rank(self):
        return len(self.edges)

    def get_edge(self, i):
        return self.edges[i]

    def get_all_edges(self):
        return list(self.edges)

    def get_all_dangling(self):
        return [e for e in self.edges if e.is_dangling()]

    def get_all_nondangling(self):
        return {e for e in self.edges if not e.is_dangling()}

    def __getitem__(self, i):
        return self.edges[i]

    def get_tensor(self):
        return self.tensor

    def set_tensor(self, t):
        self.tensor = t

    def reorder_edges(self, edge_order):
        perm = [self.edges.index(e) for e in edge_order]
        assert sorted(perm) == list(range(len(self.edges)))
        self.tensor = _transpose(self.tensor, perm)
        self.edges = list(edge_order)
        _fix_edge_axes(self)
        return self

    def reorder_axes(self, perm):
        return self.reorder_edges([self.edges[i] for i in perm])

    def copy(self, conjugate=False):
        t = _conj(self.tensor) if conjugate else self.tensor
        n = self.__class__.__new__(self.__class__)
        Node.__init__(n, t, name=self.name)
        return n

    def __matmul__(self, other):
        return contract_between(self, other)

    def __rmul__(self, lvalue):
        return Node(lvalue * self.tensor)

    # elementwise arithmetic of tn.Node (network_components.py of tensornetwork 0.4.x:
    # __add__ / __sub__ / __mul__ / __truediv__ act on the tensors, result is a fresh Node)
    @staticmethod
    def _other(o):
        return o.tensor if isinstance(o, Node) else o

    def __add__(self, other):
        return Node(self.tensor + Node._other(other))

    def __sub__(self, other):
        return Node(self.tensor - Node._other(other))

    def __mul__(self, other):
        return Node(self.tensor * Node._other(other))

    def __truediv__(self, other):
        return Node(self.tensor / Node._other(other))


class CopyNode(Node):
    pass


def _fix_edge_axes(node):
    """recompute axis indices of every edge of `node` from its edge list"""
    seen = {}
    for i, e in enumerate(node.edges):
        if e.node1 is node and e.node2 is node:  # trace edge: two slots
            if id(e) in seen:
                e.axis2 = i
            else:
                e.axis1 = i
                seen[id(e)] = True
        elif e.node1 is node:
            e.axis1 = i
        else:
            e.axis2 = i


def _transpose(t, perm):
    if hasattr(t, "permute") and not isinstance(t, np.ndarray):
        return t.permute(*perm)
    return np.transpose(t, perm)


def _conj(t):
    return t.conj() if hasattr(t, "conj") else np.conj(t)


def _tensordot(a, b, axes):
    if isinstance(a, np.ndarray) and isinstance(b, np.ndarray):
        return np.tensordot(a, b, axes)
    import torch

    return torch.tensordot(a, b, dims=axes)


def connect(e1, e2, name=None):
    assert e1.is_dangling() and e2.is_dangling(), "edges must be dangling to connect"
    n1, a1, n2, a2 = e1.node1, e1.axis1, e2.node1, e2.axis1
    new = Edge(n1, a1, n2, a2, name)
    n1.edges[a1] = new
    n2.edges[a2] = new
    return new


def get_shared_edges(a, b):
    return {e for e in a.edges if not e.is_dangling() and ((e.node1 is a and e.node2 is b) or (e.node1 is b and e.node2 is a))}


def get_all_edges(nodes):
    out = set()
    for n in nodes:
        out |= set(n.edges)
    return out


def get_subgraph_dangling(nodes):
    nodes = list(nodes)
    out = set()
    for n in nodes:
        for e in n.edges:
            if e.is_dangling() or not (any(e.node1 is m for m in nodes) and any(e.node2 is m for m in nodes)):
                out.add(e)
    return out


def contract_between(a, b, name=None, allow_outer_product=False, output_edge_order=None, axis_names=None):
    if a is b:
        return contract_trace_edges(a)
    shared = [e for e in a.edges if not e.is_dangling() and ((e.node1 is a and e.node2 is b) or (e.node1 is b and e.node2 is a))]
    # unique, keep order of appearance on a
    seen, sh = set(), []
    for e in shared:
        if id(e) not in seen:
            seen.add(id(e))
            sh.append(e)
    ax_a = [a.edges.index(e) for e in sh]
    ax_b = [b.edges.index(e) for e in sh]
    if not sh and not allow_outer_product:
        raise ValueError("No edges found between nodes and allow_outer_product=False")
    t = _tensordot(a.tensor, b.tensor, (ax_a, ax_b))
    free_a = [e for i, e in enumerate(a.edges) if i not in ax_a]
    free_b = [e for i, e in enumerate(b.edges) if i not in ax_b]
    new = Node.__new__(Node)
    new.tensor = t
    new.name = name or "__unnamed_node__"
    new.backend = None
    new.edges = free_a + free_b
    for e in new.edges:
        if e.node1 is a or e.node1 is b:
            if e.node2 is a or e.node2 is b:  # edge between a and b not in shared cannot happen; trace on result
                e.node1 = new
                e.node2 = new
            else:
                e.node1 = new
        elif e.node2 is a or e.node2 is b:
            e.node2 = new
    _fix_edge_axes(new)
    if output_edge_order is not None:
        new.reorder_edges(list(output_edge_order))
    return new


def contract_trace_edges(node):
    for e in list(node.edges):
        if e.is_trace():
            i, j = [k for k, x in enumerate(node.edges) if x is e]
            node.tensor = np.trace(node.tensor, axis1=i, axis2=j)
            node.edges = [x for x in node.edges if x is not e]
            _fix_edge_axes(node)
    return node


def contract(edge, name=None, axis_names=None):
    if edge.node1 is edge.node2:
        return contract_trace_edges(edge.node1)
    return contract_between(edge.node1, edge.node2, name=name)


def contract_parallel(edge):
    if edge.node1 is edge.node2:
        return contract_trace_edges(edge.node1)
    return contract_between(edge.node1, edge.node2)


def copy(nodes, conjugate=False):
    nodes = list(nodes)
    node_dict = {n: n.copy(conjugate=conjugate) for n in nodes}
    edge_dict = {}
    for n in nodes:
        for i, e in enumerate(n.edges):
            if e in edge_dict:
                continue
            if e.is_dangling():
                edge_dict[e] = node_dict[n].edges[i]
            elif e.node1 in node_dict and e.node2 in node_dict:
                n1, n2 = node_dict[e.node1], node_dict[e.node2]
                a1 = e.node1.edges.index(e)
                a2 = [k for k, x in enumerate(e.node2.edges) if x is e][-1]
                ne = Edge(n1, a1, n2, a2, e.name)
                n1.edges[a1] = ne
                n2.edges[a2] = ne
                edge_dict[e] = ne
            else:  # connected to a node outside the copied set: becomes dangling on the copy
                edge_dict[e] = node_dict[n].edges[i]
    return node_dict, edge_dict


def split_node(*a, **k):
    raise NotImplementedError("split_node (SVD split of two-qubit gates) is off the golden path")
