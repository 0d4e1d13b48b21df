"""Stand-in for the un-vendored `tensornetwork` dependency of the reference -- TEST / FIXTURE
INFRASTRUCTURE ONLY (oracle/make_golden.py).

The reference (tensorcircuit 0.12.1) delegates its arithmetic to tensornetwork 0.4.x, which is
not installed and not installable here.  This package provides just the API surface the
reference's statevector hot path touches -- Node / Edge wiring, copy, contract_between /
contract / contract_parallel, get_all_edges / get_subgraph_dangling, and a numpy backend base
class -- implemented from the published semantics (contract_between = tensordot over the
shared edges, result axes = a's free axes then b's; reorder_edges = transpose; copy(conjugate)
= elementwise conj).  With it the reference's *own* gate definitions, circuit wiring,
expectation graph, sampler and format conversions run unmodified, which is what the golden
fixtures in tests/golden/ are generated from.  Nothing in tensorcircuit_b200/ imports this."""

from . import backends, backend_contextmanager, network_components, network_operations  # noqa: F401
from .network_components import (  # noqa: F401
    AbstractNode, CopyNode, Edge, Node, connect, contract, contract_between, contract_parallel, get_all_edges,
    get_shared_edges, get_subgraph_dangling, copy, split_node,
)
from .backend_contextmanager import set_default_backend  # noqa: F401
from . import contractors  # noqa: F401

__version__ = "0.4.6-shim"


class FiniteMPS:  # placeholder base class (MPS simulator is off the hot path)
    pass
