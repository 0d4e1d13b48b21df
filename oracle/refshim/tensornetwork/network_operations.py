from .network_components import copy, get_all_edges, get_subgraph_dangling, get_shared_edges  # noqa: F401


def reachable(*a, **k):
    raise NotImplementedError


def get_all_nodes(edges):
    out = set()
    for e in edges:
        out.add(e.node1)
        if e.node2 is not None:
            out.add(e.node2)
    return out


def redirect_edge(*a, **k):
    raise NotImplementedError


def remove_node(*a, **k):
    raise NotImplementedError
