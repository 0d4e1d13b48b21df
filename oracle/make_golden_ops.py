"""Generate tests/golden/ref_operators.npz by running the REFERENCE's own code (see make_golden.py for
how the reference is loaded):  python oracle/make_golden_ops.py

Pins, with outputs of the reference itself:
  * quantum.PauliStringSum2COO (dense form of the COO matrix) for strings mixing X, Y, Z, and
    quantum.heisenberg_hamiltonian on a 6-site line with fields;
  * templates.measurements.operator_expectation / sparse_expectation of those operators (sparse and
    dense form) on a 6-qubit hardware-efficient circuit, complex64 and complex128;
  * the reference-held gradient values of tests/test_circuit.py:536-554 are reproduced in
    tests/test_circuit_api.py; here the gradient of a 6-qubit HEA energy is pinned by central
    differences of the reference's own energy in complex128.
"""

import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import tc_oracle as orc  # noqa: E402  (shared circuit recipes)
from oracle.ref_loader import load_reference  # noqa: E402

LS = [[1, 0, 3, 0, 0, 2], [0, 2, 2, 0, 1, 0], [3, 3, 0, 0, 0, 0], [0, 0, 0, 1, 1, 1], [2, 0, 0, 0, 0, 3], [0, 0, 3, 0, 0, 0]]
W = [0.5, 1.5, -1.0, 0.25, 0.75, -0.3]


def main():
    tc = load_reference()
    meas = importlib.import_module("tensorcircuit.templates.measurements")
    graphs = importlib.import_module("tensorcircuit.templates.graphs")
    n = 6
    out = {"ls": np.array(LS), "w": np.array(W)}
    params = np.random.default_rng(5).uniform(0, 2 * np.pi, size=[3, 2, n])
    out["params"] = params
    ops = orc.hea_circuit(n, params)
    for dt in ("complex64", "complex128"):
        tc.set_dtype(dt)
        coo = tc.quantum.PauliStringSum2COO(LS, W, numpy=True)
        out["coo_dense_" + dt] = np.asarray(coo.todense())
        hh = tc.quantum.heisenberg_hamiltonian(graphs.Line1D(n, pbc=False), hzz=1.0, hxx=0.7, hyy=0.3, hz=0.2, hx=-0.1, sparse=True, numpy=True)
        out["heis_dense_" + dt] = np.asarray(hh.todense())
        c = tc.Circuit(n)
        for name, q, p in ops:
            getattr(c, name)(*q, **p)
        out["e_coo_" + dt] = np.asarray(meas.operator_expectation(c, coo))
        out["e_coo_sparse_" + dt] = np.asarray(meas.sparse_expectation(c, coo))
        out["e_heis_" + dt] = np.asarray(meas.operator_expectation(c, hh))
        out["e_dense_" + dt] = np.asarray(meas.operator_expectation(c, np.asarray(hh.todense())))
    # gradient of the Heisenberg energy w.r.t. the HEA angles: central differences of the reference's value
    tc.set_dtype("complex128")
    hh = tc.quantum.heisenberg_hamiltonian(graphs.Line1D(n, pbc=False), hzz=1.0, hxx=0.7, hyy=0.3, hz=0.2, hx=-0.1, sparse=True, numpy=True)

    def energy(p):
        c = tc.Circuit(n)
        for name, q, kw in orc.hea_circuit(n, p):
            getattr(c, name)(*q, **kw)
        return float(np.real(meas.operator_expectation(c, hh)))

    g = np.zeros(params.size)
    h = 1e-5
    flat = params.reshape(-1)
    for k in range(params.size):
        up, dn = flat.copy(), flat.copy()
        up[k] += h
        dn[k] -= h
        g[k] = (energy(up.reshape(params.shape)) - energy(dn.reshape(params.shape))) / (2 * h)
    out["grad_heis_fd"] = g.reshape(params.shape)
    out["value_heis"] = energy(params)
    tc.set_dtype("complex64")
    path = os.path.join(ROOT, "tests", "golden", "ref_operators.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%d arrays, %.1f KiB)" % (path, len(out), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
