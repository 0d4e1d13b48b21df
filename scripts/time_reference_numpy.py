"""Time the REFERENCE's own numpy path on the config-4 recipe (build container only: it reads
/root/reference through oracle/ref_loader.py, which does not exist on the GPU box).  Follows the
reference's timing convention (tensorcircuit/utils.py:205-232 `benchmark`: one untimed staging
call, then the mean of `tries` timed calls).  The un-vendored tensornetwork / opt_einsum are the
stand-ins of oracle/refshim, so this is the reference's gate / circuit / contractor code over a
plain-numpy tensordot -- labelled "reference-numpy-over-shim".  Writes
profiles/r2_reference_numpy_cpu.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader, tc_oracle as orc  # noqa: E402


def main():
    tc = ref_loader.load_reference()
    tc.set_backend("numpy")
    tc.set_dtype("complex64")
    out = []
    for n in (14, 16, 18):
        ops = orc.random_circuit(n, 20, 3)

        def run():
            c = tc.Circuit(n)
            for name, q, p in ops:
                getattr(c, name)(*q, **p)
            return c.state()

        run()  # staging call (utils.py:216)
        tries = 2 if n >= 22 else 3
        t0 = time.perf_counter()
        for _ in range(tries):
            s = run()
        dt = (time.perf_counter() - t0) / tries
        out.append({"n": n, "gates": len(ops), "seconds": dt, "amplitude_updates_per_s": len(ops) * 2.0**n / dt, "norm": float(np.linalg.norm(s))})
        print(out[-1], flush=True)
    res = {"kind": "reference-numpy-over-shim", "where": "build container CPU (%d hardware threads visible), numpy %s" % (len(os.sched_getaffinity(0)), np.__version__),
           "convention": "tensorcircuit/utils.py:205-232: one staging call, mean of the timed calls", "recipe": "config 4: r on all + cnot matching, depth 20, complex64, wavefunction()",
           "runs": out}
    json.dump(res, open(os.path.join(ROOT, "profiles", "r2_reference_numpy_cpu.json"), "w"), indent=1)


main()
