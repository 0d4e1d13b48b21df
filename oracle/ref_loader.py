"""Load the reference's hot-path modules in this container -- TEST / FIXTURE INFRASTRUCTURE ONLY.

`import tensorcircuit` fails here (tensornetwork / opt_einsum / graphviz are not installed and
the package __init__ pulls in every subsystem).  This loader puts the stand-ins of
oracle/refshim/ on sys.path, registers an empty parent package whose __path__ is the reference
source tree (so `tensorcircuit/__init__.py` is NOT executed) and imports only the modules of
the statevector hot path: cons, backends, gates, quantum, abstractcircuit, basecircuit,
circuit.  The reference source is used where it lies, unmodified; nothing is copied."""

import importlib
import os
import sys
import types

REF = "/root/reference"
SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")


def load_reference():
    if "tensorcircuit" in sys.modules and hasattr(sys.modules["tensorcircuit"], "Circuit"):
        return sys.modules["tensorcircuit"]
    if not os.path.isdir(os.path.join(REF, "tensorcircuit")):
        raise ImportError("reference checkout not present")
    if SHIM not in sys.path:
        sys.path.insert(0, SHIM)
    pkg = types.ModuleType("tensorcircuit")
    pkg.__path__ = [os.path.join(REF, "tensorcircuit")]
    pkg.__version__ = "0.12.1"
    sys.modules["tensorcircuit"] = pkg
    for name in ("utils", "cons", "gates", "quantum", "simplify", "channels", "abstractcircuit", "basecircuit", "circuit", "densitymatrix"):
        m = importlib.import_module("tensorcircuit." + name)
        setattr(pkg, name, m)
    cons = pkg.cons
    for k in ("set_backend", "set_dtype", "set_contractor", "backend"):
        setattr(pkg, k, getattr(cons, k))
    pkg.Circuit = pkg.circuit.Circuit
    pkg.DMCircuit = pkg.densitymatrix.DMCircuit2  # circuit.expectation imports noisemodel, which wants it
    pkg.expectation = pkg.circuit.expectation
    pkg.Gate = pkg.gates.Gate
    pkg.array_to_tensor = pkg.gates.array_to_tensor
    return pkg


if __name__ == "__main__":
    tc = load_reference()
    c = tc.Circuit(2)
    c.H(0)
    c.cnot(0, 1)
    print(c.state())
