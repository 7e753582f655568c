"""Quick GPU sanity of the resident kernels (seconds): small shapes through both arithmetic policies against the oracle.
Run FIRST on a GPU visit, under `timeout -k 10 150`, so that a deadlocked persistent kernel costs minutes, not the call."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from decipher_b200 import synth                                                    # noqa: E402
from decipher_b200.capi import DeviceCatchment, KERNEL_RESIDENT, KERNEL_STREAMED, OPT_FAST_MATH   # noqa: E402
from oracle_binding import oracle_run                                              # noqa: E402
from parity import compare_flows, compare_perf                                     # noqa: E402

cases = [dict(nac=33, nriv=3, npt=2, ngp=2, nge=2, nstep=65, kW=3, seed=3, tree=True),
         dict(nac=500, nriv=3, npt=1, ngp=4, nge=1, nstep=700, kW=8, seed=synth.BASE_SEED + 2)]
for a in sys.argv[1:]:
    cases.append(dict(nac=int(a), nriv=3, npt=1, ngp=4, nge=1, nstep=300, kW=8, seed=synth.BASE_SEED + 9))
for kw in cases:
    cat = synth.synth_catchment(**kw)
    M = 37 if kw["nac"] < 100 else 600
    mc = synth.sample_parameters(cat.npt, M)
    ref = oracle_run(cat, mc.copy(order="F"), n_threads=os.cpu_count() or 8)
    dev = DeviceCatchment(cat)
    runs = (("streamed", KERNEL_STREAMED, 0), ("resident exact", KERNEL_RESIDENT, 0), ("resident fast", KERNEL_RESIDENT, OPT_FAST_MATH))
    if kw["nac"] > 512:
        runs = (("resident fast", KERNEL_RESIDENT, OPT_FAST_MATH),)
    for name, kernel, flags in runs:
        t0 = time.time()
        out = dev.run(mc.copy(order="F"), flags=flags, kernel=kernel, want_q=True)
        d = compare_flows(out["q"], ref["q"]); p = compare_perf(out["perf"], ref["perf"])
        out2 = dev.run(mc.copy(order="F"), flags=flags, kernel=kernel, want_q=False)
        assert np.array_equal(out2["perf"], out["perf"], equal_nan=True), "measures differ with / without the flow output"
        print(f"nac {kw['nac']:4d} {name:15s} ok  flows {d['max_rel']:.2e}  measures {p['max_rel']:.2e}  kernel {out['stats']['kernel_used']}  "
              f"{out['stats']['ms_physics']:.1f} ms physics  ({time.time() - t0:.1f} s)", flush=True)
    dev.close()
print("SANITY OK")
