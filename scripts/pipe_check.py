"""A/B of the pipelined gate pass (lpass_pipe_kernel) against lpass_fast_kernel on the GPU:
(1) results must be bit-identical (same rounds, same arithmetic, only the staging differs) for a
few grid shapes, (2) ms per step of the config-4 recipe with either kernel.

usage: python scripts/pipe_check.py [n_check=24] [n_time=30] [depth=20]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import tensorcircuit_b200 as tc  # noqa: E402
from tensorcircuit_b200 import engine, recipes  # noqa: E402


ENV = {0: {}, 1: {"TCB200_PIPE": "1"}, 2: {"TCB200_PIPE": "2"}, 3: {"TCB200_GATE_TMA": "1"}}  # 3: TMA-staged tiles


def run(n, depth, pipe, knobs=None, reps=1, seed=3):
    for k in ("TCB200_PIPE", "TCB200_GATE_TMA", "TCB200_PIPE_MIN_TILES", "TCB200_PIPE_GRID"):
        os.environ.pop(k, None)
    os.environ.update(ENV[int(pipe)])
    for k, v in (knobs or {}).items():
        os.environ[k] = str(v)
    c = recipes.build(tc.Circuit(n), recipes.random_circuit(n, depth, seed))
    blocks = c._fuse(c._ops, n)
    st = engine.DeviceState(n, "complex64")
    best = None
    for _ in range(reps):
        st.init_zero()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st.apply_planned(blocks)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return st, best


MODES = tuple(int(x) for x in os.environ.get("PIPE_CHECK_MODES", "1,2,3").split(","))


def main():
    n_check = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    n_time = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    depth = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    out = {"checks": [], "timing": {}}
    ref, _ = run(n_check, 8, False)
    ref_host = ref.buf.clone()
    for knobs in ({}, {"TCB200_PIPE_MIN_TILES": 1, "TCB200_PIPE_GRID": 7}, {"TCB200_PIPE_MIN_TILES": 1, "TCB200_PIPE_GRID": 1000},
                  {"TCB200_PIPE_MIN_TILES": 1, "TCB200_PIPE_GRID": 148 * 3 + 1}):
        for mode in MODES:
            st, _ = run(n_check, 8, mode, knobs)
            same = bool(torch.equal(st.buf, ref_host))
            diff = float((st.buf.view(torch.float32) - ref_host.view(torch.float32)).abs().max())
            out["checks"].append({"n": n_check, "mode": mode, "knobs": knobs, "bit_identical": same, "max_abs_diff": diff})
            print(out["checks"][-1], flush=True)
            del st
    # a small state where every CTA gets one or two tiles
    ref, _ = run(16, 6, False)
    for mode in MODES:
        st, _ = run(16, 6, mode, {"TCB200_PIPE_MIN_TILES": 1, "TCB200_PIPE_GRID": 5})
        out["checks"].append({"n": 16, "mode": mode, "bit_identical": bool(torch.equal(st.buf, ref.buf))})
        print(out["checks"][-1], flush=True)
        del st
    del ref, ref_host
    for pipe in (0,) + MODES + (0,) + MODES:
        _, ms = run(n_time, depth, pipe, reps=3)
        out["timing"].setdefault("mode%d" % pipe, []).append(ms)
        print("n=%d depth=%d pipe=%s: %.2f ms" % (n_time, depth, pipe, ms), flush=True)
    for k in ("TCB200_PIPE", "TCB200_GATE_TMA"):
        os.environ.pop(k, None)
    print(json.dumps(out))


main()
