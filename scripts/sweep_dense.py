"""Per-kernel roofline sweep (run on the GPU box): GB/s of one fused-block pass for every block
width, target placement and tile size.  Algorithmic bytes = 2 * sizeof(amp) * 2^n per launch;
CUDA-event timing on the launching stream, 2 warm-up + 5 timed launches."""

import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tensorcircuit_b200 as tc  # noqa: E402
from tensorcircuit_b200 import engine  # noqa: E402
from tensorcircuit_b200.fusion import Block  # noqa: E402


def main():
    n = int(os.environ.get("SWEEP_N", "32"))
    dtype = os.environ.get("SWEEP_DTYPE", "complex64")
    tiles = [int(x) for x in os.environ.get("SWEEP_TILES", "14,15,16").split(",")]
    ks = [int(x) for x in os.environ.get("SWEEP_KS", "1,2,3,4,5").split(",")]
    peak = 6534.8
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    st = engine.DeviceState(n, dtype)
    st.init_zero()
    rng = np.random.default_rng(0)
    amp = 8 if dtype == "complex64" else 16
    nbytes = 2.0 * amp * 2**n
    rows = []
    for tl in tiles:
        os.environ["TCB200_TILE_BYTES_LOG2"] = str(tl)
        for k in ks:
            if dtype == "complex128" and k == 5:
                continue
            places = {
                "low": tuple(range(k)),
                "high": tuple(range(n - k, n)),
                "mid": tuple(range(10, 10 + k)),
                "spread": tuple(sorted(set(np.linspace(1, n - 2, k).astype(int).tolist()))),
            }
            for name, bits in places.items():
                if len(bits) != k:
                    continue
                u = np.linalg.qr(rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k)))[0]
                blk = Block(qubits=tuple(sorted(n - 1 - b for b in bits)), bits=bits, matrix=u, batched=False, ngates=1)
                for _ in range(2):
                    st.apply_block(blk)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    st.apply_block(blk)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                gbs = nbytes / (ms * 1e-3) / 1e9
                rows.append({"tile_log2": tl, "k": k, "place": name, "ms": ms, "gbs": gbs, "frac": gbs / peak})
                print("tile=2^%d k=%d %-6s %8.3f ms %8.1f GB/s  %.3f of measured" % (tl, k, name, ms, gbs, gbs / peak), flush=True)
    os.environ.pop("TCB200_TILE_BYTES_LOG2", None)
    print("norm2", float(st.norm2()[0]))
    out = os.path.join("gpurun_out", "sweep_dense_%s_n%d.json" % (dtype, n))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"n": n, "dtype": dtype, "peak_gbs": peak, "rows": rows}, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
