"""Small driver for profiling: one dcp_run_ensemble call on a synthetic catchment."""
import argparse, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from decipher_b200 import synth
from decipher_b200.capi import *
ap = argparse.ArgumentParser()
ap.add_argument("--cat", default="C2"); ap.add_argument("--nstep", type=int, default=512)
ap.add_argument("--members", type=int, default=592); ap.add_argument("--kernel", default="resident")
ap.add_argument("--math", default="exact"); ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
cat = synth.synth_catchment(a.cat, nstep=a.nstep)
dev = DeviceCatchment(cat)
mc = synth.sample_parameters(cat.npt, a.members)
k = {"resident": KERNEL_RESIDENT, "streamed": KERNEL_STREAMED, "auto": KERNEL_AUTO}[a.kernel]
for _ in range(a.reps):
    o = dev.run(mc.copy(order="F"), flags=OPT_FAST_MATH if a.math == "fast" else 0, kernel=k, want_q=False)
    s = o["stats"]
    print(a.kernel, a.math, "units/s physics %.3e total %.3e" % (a.members*cat.nac*cat.nstep/(s["ms_physics"]*1e-3), a.members*cat.nac*cat.nstep/(s["ms_total"]*1e-3)), s)
