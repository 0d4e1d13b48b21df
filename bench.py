#!/usr/bin/env python
"""Benchmark of the statevector hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one pass of the hot path over one synthetic input: |0..0> -> every gate pass of the
seeded random circuit (r on every qubit + cnot on a random perfect matching,
depth 20) -> 10^6-shot CDF sample.  N = 1 runs n = 34 complex64 (128 GiB state, or the largest n
that fits the device, named in config.workload); N > 1 runs the distributed state (top log2 N
index bits = rank, NCCL all-to-all remaps) with n = 34 + log2 N (bench_dist.py).

metric  amplitude updates / s = recorded gates x 2^n x steps / time  (the same definition for
        this arm and for --impl reference, so the driver's ratio is a wall-clock speed-up per
        amplitude update; the reference arm runs a smaller n and says so in its config)
value   device-resident inputs (uniforms already in HBM), CUDA-event time, max over ranks
e2e     same metric through the public API (tc.Circuit gate calls -> c.sample(status=host
        array)): Python recording + fusion + planning + H2D of the uniforms from pinned memory
        + D2H of the sample indices inside the timed region
roofline  dominant kernel = lpass_fast_kernel (structure-aware gate pass: one HBM read + write of
        the state per pass): algorithmic bytes 16 * 2^n per launch / mean launch time (CUDA
        events around the gate phase of every timed step / number of launches), against the
        measured copy bandwidth; traffic = dram bytes per launch from the committed ncu capture
        at the bench size (profiles/roofline_traffic.json)
roofline_fp32  the CUDA-core bound next to it: real FMAs actually issued (the library reports
        them per pass) / gate-phase time against 148 SMs x 128 FMA/clk x max SM clock
configs  BASELINE configs 2 (28-qubit TFIM VQE energy) and 3 (vmap 1024 x 20-qubit HEA) measured
        in the same process after the headline, each with its own parity check
cpu_baseline  the C restatement of the reference's gate-by-gate CPU algorithm
        (oracle/sv_port.c, OpenMP, all host threads) on a bounded sample of the same recipe,
        plus the reference's own numpy path timed in the build container (committed fixture)
"""

import argparse
import gc
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEPTH = 20
SEED = 3
SHOTS = 10**6
METRIC = "amplitude_updates_per_s"
UNIT = "amplitude updates/s (recorded gates x 2^n / s)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=int(os.environ.get("TCB_BENCH_N", "0")), help="override qubit count (testing)")
    ap.add_argument("--depth", type=int, default=int(os.environ.get("TCB_BENCH_DEPTH", str(DEPTH))))
    ap.add_argument("--shots", type=int, default=int(os.environ.get("TCB_BENCH_SHOTS", str(SHOTS))))
    ap.add_argument("--kmax", type=int, default=int(os.environ.get("TCB_BENCH_KMAX", "0")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the config-2 / config-3 records")
    ap.add_argument("--single-block", action="store_true", help="one fused block per pass (dense_kernel only)")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def roofline_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------
def cpu_port_run(n, depth, seed, shots, budget_s=15.0):
    """Time the gate-by-gate C restatement on the same recipe at a bounded size."""
    from oracle import port, tc_oracle as orc

    ops = orc.random_circuit(n, depth, seed)
    try:  # all the host threads this process may use (torchrun sets OMP_NUM_THREADS=1 per rank)
        port.lib().svp_set_num_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    cores = int(port.lib().svp_num_threads())
    psi = np.empty(2**n, dtype=np.complex64)
    mats = []
    for name, q, p in ops:
        u = orc.gate_matrix(name, **p)
        k = len(q)
        bits = [n - 1 - x for x in q]
        order = np.argsort(bits)
        perm = list(order[::-1])
        t = np.transpose(u.reshape([2] * (2 * k)), perm + [k + x for x in perm]).reshape(2**k, 2**k)
        mats.append((sorted(bits), np.ascontiguousarray(t)))
    t0 = time.perf_counter()
    port.init_zero(psi, n)
    done = 0
    for bits, u in mats:
        port.apply(psi, n, bits, u)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    if shots and done == len(mats):
        port.sample(psi, n, np.random.default_rng(4).random(shots))
    dt = time.perf_counter() - t0
    return {"value": done * float(2**n) / dt, "gates": done, "seconds": dt, "cores": cores, "n": n}


def cpu_baseline_obj(n_cpu, depth, seed):
    r = cpu_port_run(n_cpu, depth, seed, 0)
    return {
        "value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "n_qubits": r["n"],
        "sample": "oracle/sv_port.c (OpenMP gate-by-gate restatement of cons.py:605-623; the reference itself needs tensornetwork/jax, not installable here): same random-circuit recipe at n=%d complex64, %d gates in %.1f s" % (r["n"], r["gates"], r["seconds"]),
    }


def reference_numpy_fixture():
    """The reference's own numpy path (its gates / circuit / contractor code over the tensornetwork
    stand-in of oracle/refshim), timed in the build container by scripts/time_reference_numpy.py
    with the convention of tensorcircuit/utils.py:205-232 -- /root/reference does not exist on the
    GPU box, so this is a committed measurement, not a live one."""
    p = os.path.join(ROOT, "profiles", "r2_reference_numpy_cpu.json")
    if not os.path.exists(p):
        return None
    try:
        d = json.load(open(p))
        best = max(d["runs"], key=lambda r: r["n"])
        return {"value": best["amplitude_updates_per_s"], "unit": UNIT, "kind": d["kind"], "n_qubits": best["n"], "seconds_per_run": best["seconds"],
                "cores": None, "sample": "committed fixture profiles/r2_reference_numpy_cpu.json: %s; %s" % (d["where"], d["recipe"])}
    except Exception:
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_cpu = args.n if args.n else 26
    vals = []
    t_all = time.perf_counter()
    for i in range(args.warmup + args.steps):
        budget = 4.0 if i < args.warmup else 12.0
        r = cpu_port_run(n_cpu, args.depth, SEED, 0, budget_s=budget)
        if i >= args.warmup:
            vals.append(r)
    total_gates = sum(r["gates"] for r in vals)
    total_s = sum(r["seconds"] for r in vals)
    value = total_gates * float(2**n_cpu) / total_s
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_s / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "c64", "data": "synthetic",
        "config": dict(workload_config(args.gpus, n_cpu, args.depth, 0), sample_of="the n=%d workload of the GPU arm: same recipe and seed at the largest size a %d-thread host finishes within the step budget" % (min(36, 34 + int(math.log2(args.gpus))), vals[0]["cores"])),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": vals[0]["cores"], "kind": "port",
                         "sample": "oracle/sv_port.c OpenMP gate-by-gate port of the reference algorithm; each step = the same recipe at n=%d complex64 bounded to 12 s (%d gates timed)" % (n_cpu, total_gates)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line))


def workload_config(n_gpus, n, depth, shots):
    return {
        "workload": "random circuit (r on all qubits + cnot on a random perfect matching) depth %d, n=%d complex64, wavefunction + %d-shot sample(status=...)" % (depth, n, shots),
        "n_qubits": n, "depth": depth, "shots": shots, "seed": SEED,
        "parallelism": "single GPU" if n_gpus == 1 else "state sharded over %d GPUs (top %d index bits = rank)" % (n_gpus, int(math.log2(n_gpus))),
        "l2": "inputs exceed L2 (state >> 126 MB); no flush needed",
    }


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def pick_n(requested, free_bytes, amp_bytes=8):
    if requested:
        return requested
    n = 34
    while (amp_bytes << n) > 0.9 * free_bytes and n > 20:
        n -= 1
    return n


def config2_tfim(tc, engine, recipes, torch, reps=3):
    """BASELINE config 2: 28-qubit TFIM VQE energy (2n Pauli terms via expectation_ps), complex64,
    one GPU.  Gate phase and energy phase timed with CUDA events; parity: the same energy in
    complex128 at n = 24 (size-independent property: both dtypes run the same kernels)."""
    n, layers = 28, 4
    params = np.random.default_rng(1).uniform(0, 2 * np.pi, [2 * layers, n])
    ops = recipes.tfim_vqe_circuit(n, params)
    terms = recipes.tfim_terms(n)
    pss, ws = [ps for _, ps in terms], [w for w, _ in terms]

    def gates():
        c = recipes.build(tc.Circuit(n), ops)
        c._ensure_state()
        return c

    c = gates()
    torch.cuda.synchronize()
    engine.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # gate phase = fusion (cached by structure) + pass planning + the gate passes; the Python recording of the
    # 248 gate calls is done before the clock starts (it is in `gate_phase_with_recording_ms`)
    g_ms = 0.0
    t_rec = 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        c = recipes.build(tc.Circuit(n), ops)
        torch.cuda.synchronize()
        e0.record()
        c._ensure_state()
        e1.record()
        torch.cuda.synchronize()
        t_rec += time.perf_counter() - t0
        g_ms += e0.elapsed_time(e1)
    g_ms /= reps
    g_rec_ms = 1e3 * t_rec / reps
    passes = engine.STATS["apply_launches"] // reps
    rounds = engine.STATS["gate_pass_rounds"] // reps
    fma = engine.STATS["gate_pass_fma_per_amp"] / reps

    def loop_energy():  # the reference idiom: a Python loop over c.expectation_ps
        e = 0.0
        vals = [c.expectation_ps(**_ps_kwargs(ps)) for ps in pss]
        for w, v in zip(ws, vals):
            e = e + w * v
        return float(np.real(e))

    loop_energy()
    torch.cuda.synchronize()
    engine.reset_stats()
    e0.record()
    for _ in range(reps):
        energy = loop_energy()
    e1.record()
    torch.cuda.synchronize()
    x_ms = e0.elapsed_time(e1) / reps
    xl = engine.STATS["expect_launches"] // reps
    # parity: complex64 vs complex128 at n = 24
    n2 = 24
    ops2 = recipes.tfim_vqe_circuit(n2, params[:, :n2])
    t2 = recipes.tfim_terms(n2)
    es = []
    for dt in ("complex64", "complex128"):
        tc.set_dtype(dt)
        cc = recipes.build(tc.Circuit(n2), ops2)
        es.append(float(np.real(tc.templates.measurements.pauli_sum_expectation(cc, [ps for _, ps in t2], [w for w, _ in t2]))))
        del cc
    tc.set_dtype("complex64")
    del c
    st_bytes = 8.0 * 2**n
    peak, _ = measured_peak()
    return {"config": "28-qubit TFIM VQE energy (H layer + %d x (rzz ladder, rx layer); 2n = %d Pauli strings through a loop of c.expectation_ps), complex64, 1 GPU" % (layers, len(pss)),
            "recorded_gates": len(ops), "gate_passes": passes, "gate_pass_rounds": rounds, "fma_per_amplitude": fma,
            "gate_phase_ms": g_ms, "gate_phase_with_recording_ms": g_rec_ms, "gate_phase_gbs": passes * 2 * st_bytes / (g_ms * 1e-3) / 1e9, "gate_phase_frac_of_hbm_peak": passes * 2 * st_bytes / (g_ms * 1e-3) / 1e9 / peak,
            "energy_ms": x_ms, "energy_launches": xl, "energy_reads_of_state_gbs": xl * st_bytes / (x_ms * 1e-3) / 1e9, "energy": energy,
            "parity": {"what": "same ansatz at n=24: complex64 vs complex128 energy", "c64": es[0], "c128": es[1], "rel_diff": abs(es[0] - es[1]) / max(1e-12, abs(es[1]))}}


def _ps_kwargs(ps):
    return {"x": [i for i, p in enumerate(ps) if p == 1], "y": [i for i, p in enumerate(ps) if p == 2], "z": [i for i, p in enumerate(ps) if p == 3]}


def config3_vmap(tc, engine, recipes, torch, B=1024, rank=0, world=1, reps=2):
    """BASELINE config 3: vmap batch of 1024 parameter sets on a 20-qubit HEA (rx / rzz / cnot,
    depth 4), TFIM energy per element; with world > 1 the batch is sharded over the ranks
    (tc.backend.vmap splits it; no data-path collective, one all-gather of the results)."""
    n, depth = 20, 4
    params = np.random.default_rng(2).uniform(0, 2 * np.pi, size=[B, depth, 2, n])
    terms = recipes.tfim_terms(n)
    pss, ws = [ps for _, ps in terms], [w for w, _ in terms]

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
    out = f(params)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = f(params)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps
    out = np.asarray(out)
    # parity on a subsample: the same energies one element at a time (no batch axis)
    idx = [0, B // 2, B - 1]
    single = [float(np.real(energy(params[i]))) for i in idx]
    err = max(abs(float(np.real(out[i])) - s) / max(1e-9, abs(s)) for i, s in zip(idx, single))
    return {"config": "vmap %d x 20-qubit HEA depth %d, TFIM energy (%d strings), complex64, %d GPU(s)" % (B, depth, len(pss), world),
            "wall_ms": 1e3 * wall, "states_per_s": B / wall, "energy0": float(np.real(out[0])),
            "parity": {"what": "3 batch elements vs the unbatched circuit", "max_rel_diff": err}}


def config_gradient(tc, engine, recipes, torch, n=24, layers=4, reps=2):
    """SURVEY 8(f) rank 2, next to the configs: value + gradient of the TFIM VQE energy through
    K.value_and_grad (adjoint-state sweep), wall time per call; parity: one component against a central
    difference of the engine's own value."""
    from tensorcircuit_b200 import autodiff

    terms = recipes.tfim_terms(n)
    pss, ws = [ps for _, ps in terms], [w for w, _ in terms]

    def energy(p):
        c = tc.Circuit(n)
        for i in range(n):
            c.h(i)
        for l in range(layers):
            for i in range(n - 1):
                c.rzz(i, i + 1, theta=p[2 * l, i])
            for i in range(n):
                c.rx(i, theta=p[2 * l + 1, i])
        return tc.templates.measurements.pauli_sum_expectation(c, pss, ws)

    p = np.random.default_rng(0).uniform(0, 2, size=(2 * layers, n))
    vg = tc.backend.value_and_grad(energy)
    before = dict(autodiff.ADJOINT_STATS)
    v, g = vg(p)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        v, g = vg(p)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps
    h = 1e-3
    e = np.zeros_like(p)
    e[1, 2] = h
    fd = (float(energy(p + e)) - float(energy(p - e))) / (2 * h)
    return {"config": "value_and_grad of a %d-qubit TFIM VQE energy (%d parameters, %d Pauli strings), adjoint-state sweep, complex64, 1 GPU" % (n, p.size, len(pss)),
            "wall_ms": 1e3 * wall, "value": float(v), "grad_norm": float(np.linalg.norm(g)),
            "sweeps": autodiff.ADJOINT_STATS["sweeps"] - before["sweeps"], "shift_rule_fallbacks": autodiff.ADJOINT_STATS["fallbacks"] - before["fallbacks"],
            "parity": {"what": "d/dp[1,2] vs a central difference (h = 1e-3) of the engine's own complex64 value", "gradient": float(g[1, 2]), "central_difference": fd}}


def run_ours(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import tensorcircuit_b200 as tc
    from tensorcircuit_b200 import _lib, engine, recipes
    from tensorcircuit_b200.fusion import Block

    if args.kmax:
        tc.Circuit.fusion_kmax = args.kmax
    if world > 1:
        from bench_dist import run_dist

        return run_dist(args, tc, rank, world, local)

    free, total = torch.cuda.mem_get_info()
    n = pick_n(args.n, free)
    ops = recipes.random_circuit(n, args.depth, SEED)
    ngates = len(ops)
    shots = args.shots

    # ---- device-resident leg -------------------------------------------------------------
    c0 = recipes.build(tc.Circuit(n), ops)
    use_passes = bool(tc.Circuit.use_passes) and not args.single_block
    c0.use_passes = use_passes
    blocks = c0._fuse(c0._ops, n)
    khist = {}
    for b in blocks:
        key = "%s%d" % (b.kind, len(b.bits))
        khist[key] = khist.get(key, 0) + 1
    st = engine.DeviceState(n, "complex64")
    u_host = torch.from_numpy(np.random.default_rng(4).random(shots)).pin_memory()
    u_dev = u_host.to("cuda")
    idx_dev = torch.empty(shots, dtype=torch.int64, device="cuda")
    ws = torch.empty(_lib.lib.tcb200_sample_workspace_bytes(n) + 1024, dtype=torch.uint8, device="cuda")
    stream = engine._stream()
    npass_box = [len(blocks)]

    def device_step(ev=None):
        st.init_zero()
        if ev:
            ev[0].record()
        if use_passes:
            npass_box[0] = st.apply_planned(blocks)
        else:
            st.apply_blocks(blocks)
        if ev:
            ev[1].record()
        _lib.check(_lib.lib.tcb200_sample(engine._ptr(st.buf), n, 0, engine._ptr(u_dev), shots, engine._ptr(idx_dev), None, 0.0, -1.0, engine._ptr(ws), ws.numel(), stream))

    for _ in range(args.warmup):
        device_step()
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.launch_count()
    engine.reset_stats()
    torch.cuda.synchronize()
    t_start.record()
    for i in range(args.steps):
        device_step(evs[i])
    t_end.record()
    torch.cuda.synchronize()
    launches = _lib.launch_count() - l0
    total_ms = t_start.elapsed_time(t_end)
    apply_ms = sum(a.elapsed_time(b) for a, b in evs)
    clk = clocks.stop()
    gp_rounds = engine.STATS["gate_pass_rounds"] // args.steps
    gp_free = engine.STATS["gate_pass_free_gates"] // args.steps
    gp_conf = engine.STATS["gate_pass_conflict_rounds"] // args.steps
    fma_per_amp = engine.STATS["gate_pass_fma_per_amp"] / args.steps if use_passes else float(sum(4 * 2 ** len(b.bits) for b in blocks))
    norm2 = float(st.norm2()[0])
    samples = idx_dev.cpu().numpy()
    # parity at the metric's own size: bit frequencies of the 10^6 samples against <Z_i> of the same
    # state (Z-string kernel): P(bit_i = 1) = (1 - <Z_i>) / 2 within 5 / sqrt(shots)
    zs = st.expectation_terms([0] * n, [1 << b for b in range(n)], [0] * n)[0].real
    freq = np.array([np.mean((samples >> b) & 1) for b in range(n)])
    marg_err = float(np.max(np.abs(freq - (1.0 - zs) / 2.0)))
    value = args.steps * ngates * float(2**n) / (total_ms * 1e-3)
    npass = npass_box[0]
    bytes_per_launch = 16.0 * float(2**n)
    launch_ms = apply_ms / (args.steps * npass)
    achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    traffic = roofline_traffic()
    # single fused block per pass (north-star item 2), measured live in the same run: k = 3 on
    # spread targets, 2 warm-up + 5 timed launches
    probe = {}
    for k in (3, 4):
        bits = tuple(sorted(set(np.linspace(1, n - 2, k).astype(int).tolist())))
        u = np.linalg.qr(np.random.default_rng(k).normal(size=(2**k, 2**k)) + 1j * np.random.default_rng(k + 9).normal(size=(2**k, 2**k)))[0]
        blk = Block(qubits=tuple(sorted(n - 1 - b for b in bits)), bits=bits, matrix=u, batched=False, ngates=1)
        for _ in range(2):
            st.apply_block(blk)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(5):
            st.apply_block(blk)
        p1.record()
        torch.cuda.synchronize()
        pms = p0.elapsed_time(p1) / 5
        probe["k%d" % k] = {"launch_ms": pms, "achieved": bytes_per_launch / (pms * 1e-3) / 1e9, "frac": bytes_per_launch / (pms * 1e-3) / 1e9 / peak}

    # ---- end-to-end leg through the public API ---------------------------------------------------
    del st
    gc.collect()
    torch.cuda.empty_cache()
    e2e_steps = max(1, min(args.steps, 2))

    def api_step():
        c = recipes.build(tc.Circuit(n), ops)
        c.use_passes = use_passes
        s = c.sample(batch=shots, allow_state=True, status=u_host, format="sample_int")
        del c  # refcount drop returns the 2^n buffer to the caching allocator for the next step
        if torch.cuda.memory_allocated() > (8 << n) // 2:  # never reached unless a cycle kept it alive
            gc.collect()
        return s

    s_api = api_step()  # warm-up (also allocates the state once; the allocator then reuses it)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        s_api = api_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_value = e2e_steps * ngates * float(2**n) / e2e_s
    same = bool(np.array_equal(s_api, samples))
    gc.collect()
    torch.cuda.empty_cache()

    fma_total = args.steps * float(2**n) * fma_per_amp
    fp32_peak = 148 * 128 * ((clk or {}).get("sm_max_mhz") or 1965.0) * 1e6 / 1e12
    kernel = "lpass_fast_kernel (structure-aware gate pass)" if (use_passes and engine.DeviceState.use_gate_pass) else ("cpass_kernel (staged multi-block pass)" if use_passes else "dense_kernel")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "c64", "data": "synthetic",
        "config": dict(workload_config(1, n, args.depth, shots), recorded_gates=ngates, fused_passes=npass, fused_blocks=khist, fusion_kmax=tc.Circuit.fusion_kmax,
                       gate_pass={"rounds_per_step": gp_rounds, "gates_absorbed_into_index_map": gp_free, "rounds_with_bank_conflicts": gp_conf,
                                  "fma_per_amplitude_per_step": fma_per_amp}),
        "fused_pass_updates_per_s": args.steps * npass * float(2**n) / (apply_ms * 1e-3),
        "gate_phase_ms_per_step": apply_ms / args.steps,
        "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0, "bytes_per_launch": bytes_per_launch,
                     "launch_ms": launch_ms,
                     "traffic": (bytes_per_launch * traffic["dram_bytes_per_algorithmic_byte"]) if traffic and "dram_bytes_per_algorithmic_byte" in traffic else None,
                     "traffic_source": (traffic or {}).get("source"),
                     "fp32_fma_per_amp_per_launch": fma_per_amp / max(1, npass),
                     "note": "a pass holds ~5 fused 4x4 blocks (85 FMA per amplitude): a light pass (<= 1 round) runs at 1.0 of the copy peak through this kernel, a heavy one is bound by issue slots inside its register-tile rounds (ncu: profiles/r2_lpass_phase_stalls.txt); fewer, fuller passes lower this fraction and the step time together (DESIGN.md 4.1); see roofline_fp32 and roofline_single_block"},
        "roofline_single_block": dict(probe, bound="hbm", kernel="dense_kernel", peak=peak, unit="GB/s"),
        "roofline_fp32": {"bound": "fp32 (CUDA cores)", "achieved": fma_total / (apply_ms * 1e-3) / 1e12, "peak": fp32_peak, "unit": "T FMA/s",
                          "frac": fma_total / (apply_ms * 1e-3) / 1e12 / fp32_peak, "fma_per_amplitude_per_step": fma_per_amp,
                          "note": "real FMAs issued by the gate pass (the library counts them); the same circuit gate by gate is 5440 per amplitude"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(shots * 8 + sum(16 * 4 ** len(b.bits) for b in blocks)), "d2h_bytes_per_step": int(shots * 8),
                "ms_per_step": 1e3 * e2e_s / e2e_steps, "steps": e2e_steps, "samples_match_device_leg": same},
        "gpu_launches": int(launches),
        "clocks": clk,
        "checks": {"norm2": norm2, "sample_min": int(samples.min()), "sample_max": int(samples.max()),
                   "marginals_vs_expectation_z_max_abs_diff": marg_err, "marginals_tolerance": 5.0 / math.sqrt(shots)},
    }
    if not args.no_configs:
        cfgs = []
        for fn in (config2_tfim, config3_vmap, config_gradient):
            try:
                cfgs.append(fn(tc, engine, recipes, torch))
            except Exception as e:  # extra records never cost the headline number
                cfgs.append({"config": fn.__name__, "failed": repr(e)})
            gc.collect()
            torch.cuda.empty_cache()
        line["configs"] = cfgs
    if not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline_obj(26, args.depth, SEED)
        except Exception as e:  # the baseline is a report, never a reason to lose the GPU number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": "failed: %r" % (e,)}
        ref_np = reference_numpy_fixture()
        if ref_np:
            line["cpu_baseline_reference_numpy"] = ref_np
    print(json.dumps(line))


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
