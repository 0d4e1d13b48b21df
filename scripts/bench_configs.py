"""Timings of BASELINE.json configs 2 and 3 (not bench.py lines; reported in profiles/README.md).

config 2: 28-qubit TFIM VQE energy (56 Pauli strings) complex64 -- gate phase and expectation phase
config 3: vmap batch of 1024 parameter sets on a 20-qubit HEA (TFIM energy), one GPU
CUDA-event timing, 1 warm-up + 3 timed repetitions."""

import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tensorcircuit_b200 as tc  # noqa: E402
from tensorcircuit_b200 import engine, recipes  # noqa: E402


def timed(f, reps=3):
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        out = f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (time.perf_counter() - t0) * 1e3 / reps, out


def config2():
    n = 28
    params = np.random.default_rng(1).uniform(0, 2 * np.pi, [8, n])
    ops = recipes.tfim_vqe_circuit(n, params)
    terms = recipes.tfim_terms(n)
    pss = [ps for _, ps in terms]
    ws = [w for w, _ in terms]

    def gates():
        c = recipes.build(tc.Circuit(n), ops)
        c._ensure_state()
        return c

    engine.reset_stats()
    g_ms, g_wall, c = timed(gates)
    launches = engine.STATS["apply_launches"] // 4

    def expect():
        return tc.templates.measurements.pauli_sum_expectation(c, pss, ws)

    engine.reset_stats()
    e_ms, e_wall, energy = timed(expect)
    elaunch = engine.STATS["expect_launches"] // 4
    st_bytes = 8.0 * 2**n
    return {
        "config": "28-qubit TFIM VQE energy, 56 Pauli strings, complex64", "recorded_gates": len(ops), "gate_passes": launches,
        "gate_phase_ms": g_ms, "gate_phase_wall_ms": g_wall, "gate_pass_gbs": launches * 2 * st_bytes / (g_ms * 1e-3) / 1e9,
        "expect_launches": elaunch, "expect_ms": e_ms, "expect_wall_ms": e_wall, "expect_gbs": elaunch * st_bytes / (e_ms * 1e-3) / 1e9,
        "energy": float(energy),
    }


def config3():
    n, B, depth = 20, 1024, 4
    params = np.random.default_rng(2).uniform(0, 2 * np.pi, size=[B, depth, 2, n])
    terms = recipes.tfim_terms(n)
    pss = [ps for _, ps in terms]
    ws = [w for w, _ in terms]

    def energy(p):
        c = tc.Circuit(n)
        for l in range(depth):
            for i in range(n):
                c.rx(i, theta=p[l, 0, i])
            for i in range(n - 1):
                c.rzz(i, i + 1, theta=p[l, 1, i])
            for i in range(n - 1):
                c.cnot(i, i + 1)
        return tc.templates.measurements.pauli_sum_expectation(c, pss, ws)

    f = tc.backend.vmap(energy)
    engine.reset_stats()
    ms, wall, out = timed(lambda: f(params), reps=2)
    launches = engine.STATS["apply_launches"] // 3
    st_bytes = 8.0 * 2**n * B
    return {
        "config": "vmap 1024 x 20-qubit HEA depth 4, TFIM energy (40 strings), complex64, 1 GPU", "recorded_gates_per_element": depth * (3 * n - 2),
        "gate_passes": launches, "total_ms": ms, "wall_ms": wall, "states_per_s": B / (wall * 1e-3),
        "gate_pass_gbs_if_all_gate_time": launches * 2 * st_bytes / (ms * 1e-3) / 1e9, "energy0": float(out[0]),
    }


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "3":  # config 3 only (for ncu launch lists)
        print(json.dumps(config3(), indent=1))
        sys.exit(0)
    res = {"config2": config2(), "config3": config3()}
    print(json.dumps(res, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/bench_configs.json", "w"), indent=1)
