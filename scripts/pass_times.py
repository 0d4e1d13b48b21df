"""Per-pass timing of a recipe through the planner (GPU): rounds, FMA per amplitude, gathered
bits and milliseconds of every gate pass -- the data the pass-cost model is fitted on."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import tensorcircuit_b200 as tc  # noqa: E402
from tensorcircuit_b200 import engine, fusion, recipes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 20
c = recipes.build(tc.Circuit(n), recipes.random_circuit(n, depth, 3))
blocks = c._fuse(c._ops, n)
st = engine.DeviceState(n, "complex64")
st.init_zero()
rows = []
orig = engine.DeviceState.apply_gate_pass


def timed(self, blks, tile_hi):
    b0 = dict(engine.STATS)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = orig(self, blks, tile_hi)
    e1.record()
    torch.cuda.synchronize()
    rows.append({"blocks": len(blks), "hi": len(tile_hi), "rounds": engine.STATS["gate_pass_rounds"] - b0["gate_pass_rounds"],
                 "fma": engine.STATS["gate_pass_fma_per_amp"] - b0["gate_pass_fma_per_amp"], "ms": e0.elapsed_time(e1)})
    return r


for rep in range(2):
    rows.clear()
    st.init_zero()
    engine.DeviceState.apply_gate_pass = timed
    st.apply_planned(blocks)
    engine.DeviceState.apply_gate_pass = orig
print(json.dumps({"n": n, "passes": rows, "total_ms": sum(r["ms"] for r in rows)}))
