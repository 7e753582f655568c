import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from decipher_b200 import synth
from decipher_b200.capi import *
from oracle_binding import oracle_run
cat = synth.synth_catchment("C2", nstep=3000)
M = 256
mc = synth.sample_parameters(1, M)
ref = oracle_run(cat, mc.copy(order='F'), n_threads=16, want_totals=False)
dev = DeviceCatchment(cat)
import os
cases = (("lean", OPT_FAST_MATH, KERNEL_RESIDENT),) if os.environ.get("LEAN_ONLY") else (("exact", 0, KERNEL_RESIDENT), ("lean", OPT_FAST_MATH, KERNEL_RESIDENT), ("fast-streamed", OPT_FAST_MATH, KERNEL_STREAMED))
for name, flags, k in cases:
    out = dev.run(mc.copy(order='F'), flags=flags, kernel=k, want_q=True)
    a, b = out['q'], ref['q']
    fin = np.isfinite(b).all(axis=(0,1))
    scale = np.abs(b).mean(axis=1, keepdims=True)
    err = np.abs(a-b)/np.maximum(np.abs(b), 1e-3*scale+1e-300)
    err = err[:, :, fin]
    pm = err.max(axis=(0,1))
    print(name, "finite", fin.sum(), "max", pm.max(), "median member max", np.median(pm), "99pct", np.quantile(pm, 0.99), "worst members", np.argsort(-pm)[:5], np.sort(pm)[-5:])
    pe = np.abs(1-out['perf'][:, :, fin]) ; pr = np.abs(1-ref['perf'][:, :, fin])
    print("   perf max rel", np.nanmax(np.abs(pe-pr)/pr))
