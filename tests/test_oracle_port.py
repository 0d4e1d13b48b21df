"""The C restatement (oracle/sv_port.c, used as the CPU baseline) agrees with the numpy oracle."""

import numpy as np
import pytest

from oracle import port
from oracle import tc_oracle as orc


@pytest.mark.parametrize("dtype,tol", [(np.complex64, 2e-6), (np.complex128, 1e-13)])
def test_port_matches_oracle(dtype, tol):
    n = 11
    ops = orc.random_circuit(n, 4, seed=3) + orc.hea_circuit(n, np.random.default_rng(0).uniform(0, 6, size=[2, 2, n]))
    ops.append(("toffoli", (7, 2, 9), {}))
    ops.append(("iswap", (10, 0), {"theta": 0.3}))
    psi = port.run_gatelist(n, ops, dtype)
    ref = orc.run_gatelist(n, ops).state()
    assert np.linalg.norm(psi - ref) / np.linalg.norm(ref) < 50 * tol
    rng = np.random.default_rng(1)
    for ps in rng.integers(0, 4, size=[6, n]):
        x, y, z = orc.resolve_ps(n, ps=list(ps))
        f, s, ny = orc.pauli_masks(n, x, y, z)
        want = orc.pauli_expectation(psi.astype(np.complex128), n, x, y, z)
        assert abs(port.expect(psi, n, f, s, ny) - want) < 1e-10
    u = rng.random(5000)
    got = port.sample(psi, n, u)
    want = orc.probability_sample(np.abs(psi.astype(np.complex128)) ** 2, u)
    assert np.mean(got == want) > 0.999
    assert port.lib().svp_num_threads() >= 1


def test_recipes_match_oracle():
    """the product-side benchmark recipes are the oracle's recipes"""
    from tensorcircuit_b200 import recipes

    assert recipes.random_circuit(7, 3, 5) == orc.random_circuit(7, 3, 5)
    p = np.random.default_rng(0).uniform(0, 6, size=[2, 2, 5])
    assert recipes.hea_circuit(5, p) == orc.hea_circuit(5, p)
    q = np.random.default_rng(0).uniform(0, 6, size=[4, 5])
    assert recipes.tfim_vqe_circuit(5, q) == orc.tfim_vqe_circuit(5, q)
    assert recipes.tfim_terms(6) == orc.tfim_terms(6)
