"""Multi-GPU leg of bench.py (config 5): the random-circuit recipe on a state sharded over N
GPUs -- top log2(N) index bits = rank, NCCL all-to-all global<->local qubit remaps -- then a
10^6-shot distributed CDF sample.  Launched by torchrun, one rank per GPU.

n = min(36, 34 + log2 N): 35 qubits on 2 GPUs (2^36 complex64 = 512 GiB does not fit on two
180 GB devices), 36 on 4 (128 GiB shards, chunked exchange) and on 8 (64 GiB shards, double
buffered).  Timed on the device, barrier + synchronize on both sides, max over ranks."""

import json
import math
import time

import numpy as np
import torch
import torch.distributed as dist


def run_dist(args, tc, rank, world, local):
    from bench import METRIC, UNIT, SEED, ClockSampler, measured_peak, workload_config
    from tensorcircuit_b200 import _lib, recipes
    from tensorcircuit_b200 import engine
    from tensorcircuit_b200.dist import DistState

    g = int(round(math.log2(world)))
    n = args.n if args.n else min(36, 34 + g)
    ops = recipes.random_circuit(n, args.depth, SEED)
    ngates = len(ops)
    c0 = recipes.build(tc.Circuit(n), ops)
    blocks = c0._fuse(c0._ops, n)

    # ---- parity of the distributed path before anything is timed: the same recipe at n = 22 on
    # the sharded state against the oracle (rank 0 holds the oracle; every rank checks its shard)
    from oracle import tc_oracle as orc

    n_par = 22
    ops_par = recipes.random_circuit(n_par, 6, SEED)
    c_par = recipes.build(tc.Circuit(n_par), ops_par)
    ds_par = DistState(n_par, "complex64")
    ds_par.init_zero()
    ds_par.run(c_par._fuse(c_par._ops, n_par))
    got = ds_par.gather_state()
    ref = orc.run_gatelist(n_par, ops_par).state()
    dist_parity = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
    del ds_par, got, ref, c_par
    shots = args.shots
    u = np.random.default_rng(4).random(shots)
    u_host = torch.from_numpy(u).pin_memory()
    u_dev = u_host.to("cuda")

    ds = DistState(n, "complex64")

    def step():
        ds.init_zero()
        ds.run(blocks)
        return ds.sample(u_dev)

    def sync():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync()
    ds.finalize_stats()
    for k in ds.stats:
        ds.stats[k] = 0
    engine.reset_stats()
    clocks = ClockSampler(local) if rank == 0 else None
    l0 = _lib.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    t0.record()
    for _ in range(args.steps):
        s = step()
    t1.record()
    sync()
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    launches = _lib.launch_count() - l0
    norm2 = ds.norm2()
    stats = dict(ds.finalize_stats())
    gp_stats = {"rounds_per_step": engine.STATS["gate_pass_rounds"] / max(1, args.steps), "fma_per_amplitude_per_step": engine.STATS["gate_pass_fma_per_amp"] / max(1, args.steps),
                "gates_absorbed_into_index_map_per_step": engine.STATS["gate_pass_free_gates"] / max(1, args.steps)}
    clk = clocks.stop() if clocks else None

    # end-to-end through the public API (SPMD: every rank records the circuit, state sharded)
    del ds
    import gc

    gc.collect()
    torch.cuda.empty_cache()
    tc.set_distributed(True)
    e2e_steps = 1

    def api_step(keep=False):
        c = recipes.build(tc.Circuit(n), ops)
        r = c.sample(batch=shots, allow_state=True, status=u_host, format="sample_int")
        if keep:
            return r, c
        del c
        gc.collect()
        return r, None

    api_step()
    sync()
    w0 = time.perf_counter()
    for k in range(e2e_steps):
        s_api, c_last = api_step(keep=(k == e2e_steps - 1))
    sync()
    e2e_s = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device="cuda")
    dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    same = bool(np.array_equal(s_api, s))

    # config 5's second half: all 2n TFIM strings on the sharded state (not part of the timed steps)
    terms = recipes.tfim_terms(n)
    sync()
    w1 = time.perf_counter()
    energy = float(np.real(tc.templates.measurements.pauli_sum_expectation(c_last, [ps for _, ps in terms], [w for w, _ in terms])))
    sync()
    exp_s = torch.tensor([time.perf_counter() - w1], dtype=torch.float64, device="cuda")
    dist.all_reduce(exp_s, op=dist.ReduceOp.MAX)
    exp_s = float(exp_s.item())
    del c_last
    gc.collect()

    # BASELINE config 3 with the batch sharded over the ranks (north star: batched config at 2/4/8 GPUs)
    cfg3 = None
    if not getattr(args, "no_configs", False):
        try:
            tc.set_distributed(False)
            from bench import config3_vmap

            sync()  # backend.vmap splits the batch over the ranks of the default group by itself
            cfg3 = config3_vmap(tc, engine, recipes, torch, B=1024, rank=rank, world=world)
            sync()
            w = torch.tensor([cfg3["wall_ms"]], dtype=torch.float64, device="cuda")
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
            cfg3["wall_ms"] = float(w.item())
            cfg3["states_per_s"] = 1024 / (cfg3["wall_ms"] * 1e-3)
        except Exception as e:
            cfg3 = {"config": "config3_vmap sharded", "failed": repr(e)}

    if rank == 0:
        value = args.steps * ngates * float(2**n) / (total_ms * 1e-3)
        peak, peak_src = measured_peak()
        shard_bytes = 8.0 * 2 ** (n - g)
        local_ms = (total_ms - stats["remap_ms"]) / max(1, stats["local_passes"] + stats["swap_passes"])
        achieved = 2 * shard_bytes / (local_ms * 1e-3) / 1e9
        remap_gbs = (stats["remap_bytes"] / max(1e-9, stats["remap_ms"] * 1e-3)) / 1e9 if stats["remaps"] else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "c64", "data": "synthetic",
            "config": dict(workload_config(world, n, args.depth, shots), recorded_gates=ngates, fused_blocks=len(blocks),
                           shard_gib=shard_bytes / 2**30, exchange="double-buffered all_to_all" if shard_bytes * 2 < 150 * 2**30 else "chunked all_to_all through staging"),
            "roofline": {"bound": "hbm", "kernel": "lpass_fast_kernel (structure-aware gate passes on the shard between remaps)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src, "traffic": None},
            "remap": {"per_step": stats["remaps"] / args.steps, "bytes_per_rank_per_remap": (stats["remap_bytes"] / stats["remaps"]) if stats["remaps"] else 0,
                      "ms_per_remap": (stats["remap_ms"] / stats["remaps"]) if stats["remaps"] else 0, "nvlink_gbs_per_direction": remap_gbs,
                      "nvlink_peak_gbs": 900.0, "nvlink_peak_source": "nominal NVLink 5 per direction per GPU (B200_PROFILING.md; measured peer-copy reference on this pool: 770)",
                      "nvlink_frac": (remap_gbs / 900.0) if remap_gbs else None,
                      "local_passes_per_step": stats["local_passes"] / args.steps, "swap_passes_per_step": stats["swap_passes"] / args.steps,
                      "swap_blocks_folded_into_passes_per_step": stats.get("swap_blocks", 0) / args.steps,
                      "remap_share_of_step": stats["remap_ms"] / total_ms},
            "e2e": {"value": e2e_steps * ngates * float(2**n) / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(shots * 8), "d2h_bytes_per_step": int(shots * 8),
                    "ms_per_step": 1e3 * e2e_s / e2e_steps, "steps": e2e_steps, "samples_match_device_leg": same},
            "expectation": {"strings": len(terms), "what": "TFIM energy, X_i and Z_i Z_i+1 on the sharded state (remaps for global X included)",
                            "ms": 1e3 * exp_s, "energy": energy},
            "gpu_launches": int(launches),
            "clocks": clk,
            "checks": {"norm2": norm2, "sample_min": int(s.min()), "sample_max": int(s.max()),
                       "dist_parity_relerr": dist_parity, "dist_parity_what": "config-5 recipe at n=22 depth 6 on the sharded state (remaps included) vs the oracle, relative l2 error; tolerance 1e-5"},
            "gate_pass": gp_stats,
        }
        if cfg3 is not None:
            line["configs"] = [cfg3]
        print(json.dumps(line))
