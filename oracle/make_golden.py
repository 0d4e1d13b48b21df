"""Generate tests/golden/ref_fixtures.npz by running the REFERENCE's own code.

Run in the build container (needs /root/reference):  python oracle/make_golden.py

The reference's hot-path modules (gates.py, abstractcircuit.py, basecircuit.py, circuit.py,
quantum.py, backends/abstract_backend.py + numpy_backend.py, cons.py contractor) execute
unmodified from /root/reference through oracle/ref_loader.py; only their un-vendored
third-party dependency (tensornetwork / opt_einsum / graphviz) is replaced by the stand-ins in
oracle/refshim/.  The fixtures pin, with outputs of the reference itself:
  * amplitudes of BASELINE config 1 (10-qubit HEA) in complex64 and complex128, and of the
    config-4 random-circuit recipe at n = 9, and of a circuit that uses every gate name;
  * expectation_ps of the TFIM strings + random strings, and expectation of general operators;
  * sample(allow_state=True, status=u) indices -- the one quantity no reference test pins;
  * sample formats and numpy-backend vmap values;
  * readout error: noisy probabilities, samples and sample_expectation_ps;
  * measure / perfect_sampling outcomes and record probabilities for given per-qubit status;
  * a 14-qubit QAOA MaxCut circuit: strided amplitudes, every ZZ cost term, samples;
  * Monte-Carlo noise trajectories (depolarizing / amplitudedamping / phasedamping / reset /
    cond_measure / unitary_kraus / mid_measurement) driven by fixed ``status`` values.
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import tc_oracle as orc  # noqa: E402  (only for the shared circuit recipes)
from oracle.ref_loader import load_reference  # noqa: E402


def build(tc, n, ops, **kw):
    c = tc.Circuit(n, **kw)
    for name, q, p in ops:
        getattr(c, name)(*q, **p)
    return c


ALL_GATES = [
    ("h", (0,), {}), ("x", (1,), {}), ("y", (2,), {}), ("z", (3,), {}), ("t", (0,), {}), ("s", (1,), {}), ("td", (2,), {}), ("sd", (3,), {}),
    ("wroot", (0,), {}), ("cnot", (0, 1), {}), ("cz", (1, 2), {}), ("swap", (2, 3), {}), ("cy", (3, 0), {}), ("ox", (0, 2), {}), ("oy", (1, 3), {}),
    ("oz", (2, 0), {}), ("toffoli", (0, 1, 2), {}), ("fredkin", (3, 1, 0), {}),
    ("r", (0,), dict(theta=0.3, alpha=1.1, phi=-0.7)), ("cr", (1, 2), dict(theta=0.3, alpha=1.1, phi=-0.7)), ("u", (3,), dict(theta=0.4, phi=0.5, lbd=-1.2)),
    ("cu", (0, 3), dict(theta=0.4, phi=0.5, lbd=-1.2)), ("rx", (1,), dict(theta=0.6)), ("ry", (2,), dict(theta=0.7)), ("rz", (3,), dict(theta=0.8)),
    ("phase", (0,), dict(theta=0.9)), ("rxx", (0, 1), dict(theta=0.25)), ("ryy", (1, 2), dict(theta=0.35)), ("rzz", (2, 3), dict(theta=0.45)),
    ("cphase", (3, 1), dict(theta=0.55)), ("crx", (0, 2), dict(theta=0.65)), ("cry", (1, 3), dict(theta=0.75)), ("crz", (2, 0), dict(theta=0.85)),
    ("orx", (3, 2), dict(theta=0.15)), ("ory", (0, 1), dict(theta=0.22)), ("orz", (1, 0), dict(theta=0.33)), ("iswap", (2, 3), dict(theta=0.4)),
    ("cx", (3, 2), {}), ("cswap", (0, 2, 3), {}), ("ccnot", (1, 2, 0), {}), ("sdg", (1,), {}), ("tdg", (2,), {}),
]


def main():
    tc = load_reference()
    out = {}
    # ---- config 1: 10-qubit HEA, both dtypes -------------------------------------------------
    n = 10
    params = np.random.default_rng(0).uniform(0, 2 * np.pi, size=[4, 2, n])
    ops = orc.hea_circuit(n, params)
    pss = [ps for _, ps in orc.tfim_terms(n)] + [list(r) for r in np.random.default_rng(0).integers(0, 4, size=[8, n])]
    out["hea10_pss"] = np.array(pss)
    for dt in ("complex64", "complex128"):
        tc.set_dtype(dt)
        c = build(tc, n, ops)
        out["hea10_state_" + dt] = np.asarray(c.wavefunction())
        out["hea10_exps_" + dt] = np.array([np.asarray(c.expectation_ps(ps=ps)) for ps in pss])
    tc.set_dtype("complex64")
    # ---- config 4 recipe at n = 9: amplitudes + sample(status=...) --------------------------------
    n = 9
    ops = orc.random_circuit(n, 6, seed=3)
    c = build(tc, n, ops)
    out["rand9_state"] = np.asarray(c.wavefunction())
    u = np.random.default_rng(4).random(2000)
    out["rand9_status"] = u
    out["rand9_sample_int"] = np.asarray(c.sample(batch=2000, allow_state=True, status=u, format="sample_int"))
    out["rand9_sample_bin_head"] = np.asarray(c.sample(batch=8, allow_state=True, status=u[:8], format="sample_bin"))
    out["rand9_count_vector"] = np.asarray(c.sample(batch=2000, allow_state=True, status=u, format="count_vector"))
    tc.set_dtype("complex128")
    c = build(tc, n, ops)
    out["rand9_sample_int_c128"] = np.asarray(c.sample(batch=2000, allow_state=True, status=u, format="sample_int"))
    out["rand9_probability_c128"] = np.asarray(c.probability())
    tc.set_dtype("complex64")
    # ---- every gate name -----------------------------------------------------------------------
    c = build(tc, 4, ALL_GATES)
    out["allgates_state"] = np.asarray(c.wavefunction())
    inp = np.random.default_rng(7).normal(size=16) + 1j * np.random.default_rng(8).normal(size=16)
    inp /= np.linalg.norm(inp)
    out["allgates_inputs"] = inp
    c = build(tc, 4, ALL_GATES, inputs=inp)
    out["allgates_state_inputs"] = np.asarray(c.wavefunction())
    out["allgates_exp_x0z2"] = np.asarray(c.expectation_ps(x=[0], z=[2]))
    out["allgates_exp_y1y3"] = np.asarray(c.expectation_ps(y=[1, 3]))
    op2 = np.random.default_rng(9).normal(size=(4, 4)) + 1j * np.random.default_rng(10).normal(size=(4, 4))
    out["allgates_op2"] = op2
    out["allgates_exp_op2_q12"] = np.asarray(c.expectation((op2.reshape(2, 2, 2, 2).astype(np.complex64), [1, 2]), (tc.gates.z(), [0])))
    # complex gate parameter (non-unitary), tests/test_circuit.py:343-352
    c = tc.Circuit(2)
    c.rx(0, theta=0.8 + 0.7j)
    c.rzz(0, 1, theta=-0.2j)
    out["complex_param_state"] = np.asarray(c.wavefunction())
    # unitary-as-input form, tests/test_circuit.py:404-412
    c = tc.Circuit(2, inputs=np.eye(4))
    c.rx(0, theta=0.3)
    c.cnot(0, 1)
    out["unitary_inputs_state"] = np.asarray(c.wavefunction())
    # ---- numpy-backend vmap values (numpy_backend.py:394-418) --------------------------------------
    K = tc.backend

    def f(theta):
        c = tc.Circuit(3)
        c.rx(0, theta=theta[0])
        c.ry(1, theta=theta[1])
        c.cnot(0, 2)
        c.rzz(1, 2, theta=theta[2] * 0.5)
        return K.real(c.expectation_ps(z=[2]) + 2.0 * c.expectation_ps(x=[1], z=[0]))

    th = np.random.default_rng(0).uniform(0, 2, size=(5, 3))
    out["vmap_theta"] = th
    out["vmap_values"] = np.asarray(K.vmap(f, vectorized_argnums=0)(th))
    # ---- Monte-Carlo trajectories with fixed status (circuit.py:473-744, basecircuit.py:824-857) ---
    n = 5
    rng = np.random.default_rng(21)
    ntraj = 12
    st = rng.random((ntraj, orc.NOISY_STATUS_LEN(n)))
    th = rng.uniform(0, 2 * np.pi, size=(ntraj, 2 * n))
    out["traj_status"] = st
    out["traj_theta"] = th
    for dt in ("complex64", "complex128"):
        tc.set_dtype(dt)
        states, picks = [], []
        for t in range(ntraj):
            c = tc.Circuit(n)
            picks.append(orc.noisy_trajectory(c, n, st[t], th[t]))
            states.append(np.asarray(c.wavefunction()))
        out["traj_states_" + dt] = np.array(states)
        out["traj_picks_" + dt] = np.array(picks)
    tc.set_dtype("complex64")
    # ---- sample_expectation_ps (basecircuit.py:618-758): exact and from shots with status -----------
    c = build(tc, 4, ALL_GATES)
    out["sexpps_exact"] = np.array([np.asarray(c.sample_expectation_ps(x=[0], y=[1], z=[3])),
                                    np.asarray(c.sample_expectation_ps(z=[0, 2])),
                                    np.asarray(c.sample_expectation_ps(y=[2, 3]))])
    u = np.random.default_rng(12).random(4096)
    out["sexpps_status"] = u
    out["sexpps_shots"] = np.array([np.asarray(c.sample_expectation_ps(x=[0], y=[1], z=[3], shots=4096, status=u)),
                                    np.asarray(c.sample_expectation_ps(x=[1, 2], shots=4096, status=u))])
    # ---- readout error (basecircuit.py:587-596, 760-803) ------------------------------------------------
    ro = [[0.9, 0.75], [0.4, 0.7], [0.95, 0.97], [0.6, 0.8]]
    out["readout_error"] = np.array(ro)
    c = build(tc, 4, ALL_GATES)
    out["readout_probs"] = np.asarray(c.readouterror_bs(ro, c.probability()))
    u = np.random.default_rng(14).random(2000)
    out["readout_status"] = u
    out["readout_sample_int"] = np.asarray(c.sample(batch=2000, allow_state=True, readout_error=ro, status=u, format="sample_int"))
    out["readout_sexpps"] = np.array([np.asarray(c.sample_expectation_ps(x=[0], y=[1], z=[3], readout_error=ro)),
                                      np.asarray(c.sample_expectation_ps(z=[0, 2], readout_error=ro)),
                                      np.asarray(c.sample_expectation_ps(x=[1, 2], shots=2000, status=u, readout_error=ro))])
    # ---- measure_jit / perfect_sampling with status (basecircuit.py:359-443) ---------------------------
    st = np.random.default_rng(13).random((16, 4))
    out["measure_status"] = st
    for dt in ("complex64", "complex128"):
        tc.set_dtype(dt)
        c = build(tc, 4, ALL_GATES)
        bits, probs, pbits, pprobs = [], [], [], []
        for row in st:
            b, pr = c.measure(0, 2, 3, with_prob=True, status=row)
            bits.append(np.asarray(b))
            probs.append(np.asarray(pr))
            b, pr = c.perfect_sampling(status=row)
            pbits.append(np.asarray(b))
            pprobs.append(np.asarray(pr))
        out["measure_bits_" + dt] = np.array(bits)
        out["measure_probs_" + dt] = np.array(probs)
        out["perfect_bits_" + dt] = np.array(pbits)
        out["perfect_probs_" + dt] = np.array(pprobs)
    tc.set_dtype("complex64")
    # ---- QAOA MaxCut, n = 14, p = 2 (diagonal cost function: every term is a ZZ string) -------------
    n = 14
    edges = [(i, (i + 1) % n) for i in range(n)] + [(i, (i + 5) % n) for i in range(0, n, 2)]
    gam, bet = [0.37, -0.61], [0.52, 0.23]
    out["qaoa_edges"] = np.array(edges)
    out["qaoa_angles"] = np.array([gam, bet])
    for dt in ("complex64", "complex128"):
        tc.set_dtype(dt)
        c = tc.Circuit(n)
        for i in range(n):
            c.h(i)
        for l in range(2):
            for a, b in edges:
                c.rzz(a, b, theta=2 * gam[l])
            for i in range(n):
                c.rx(i, theta=2 * bet[l])
        psi = np.asarray(c.wavefunction())
        out["qaoa_amps_" + dt] = psi[:: 2**n // 256]          # 256 strided amplitudes
        out["qaoa_zz_" + dt] = np.array([np.asarray(c.expectation_ps(z=[a, b])) for a, b in edges])
        out["qaoa_z3_" + dt] = np.asarray(c.expectation_ps(z=[0, 6, 13]))
        if dt == "complex64":
            u = np.random.default_rng(11).random(512)
            out["qaoa_status"] = u
            out["qaoa_sample_int"] = np.asarray(c.sample(batch=512, allow_state=True, status=u, format="sample_int"))
    tc.set_dtype("complex64")
    path = os.path.join(ROOT, "tests", "golden", "ref_fixtures.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%d arrays, %.1f KiB)" % (path, len(out), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
