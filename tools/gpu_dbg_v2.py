"""Debug aid: the FAST resident kernel alone on two small shapes (used with DCP_B200_LIB=<variant build> under a short timeout)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from decipher_b200 import synth
from decipher_b200.capi import DeviceCatchment, KERNEL_RESIDENT, OPT_FAST_MATH
from oracle_binding import oracle_run
from parity import compare_flows, compare_perf
which = sys.argv[1] if len(sys.argv) > 1 else "small"
kw = dict(nac=33, nriv=3, npt=2, ngp=2, nge=2, nstep=65, kW=3, seed=3, tree=True) if which == "small" else \
     dict(nac=500, nriv=3, npt=1, ngp=4, nge=1, nstep=int(sys.argv[2]) if len(sys.argv) > 2 else 200, kW=8, seed=synth.BASE_SEED + 2)
cat = synth.synth_catchment(**kw)
M = int(os.environ.get("DBG_M", "37" if which == "small" else "600"))
mc = synth.sample_parameters(cat.npt, M)
ref = oracle_run(cat, mc.copy(order="F"), n_threads=os.cpu_count() or 8)
dev = DeviceCatchment(cat)
t0 = time.time()
out = dev.run(mc.copy(order="F"), flags=OPT_FAST_MATH, kernel=KERNEL_RESIDENT, want_q=True)
print(which, "ran in %.2f s, physics %.1f ms" % (time.time() - t0, out["stats"]["ms_physics"]), flush=True)
print(compare_flows(out["q"], ref["q"])["max_rel"], compare_perf(out["perf"], ref["perf"])["max_rel"], flush=True)
