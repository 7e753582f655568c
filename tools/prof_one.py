"""One resident fast-math launch for ncu: C2 catchment, QV_M members (default 592 = one wave), QV_NSTEP steps (default 512)."""
import os, sys
sys.path.insert(0, '/root/repo')
from decipher_b200 import synth
from decipher_b200.capi import *
nstep = int(os.environ.get("QV_NSTEP", "512")); M = int(os.environ.get("QV_M", "592"))
cat = synth.synth_catchment("C2", nstep=nstep)
dev = DeviceCatchment(cat)
mc = synth.sample_parameters(1, M)
o = dev.run(mc.copy(order='F'), flags=OPT_FAST_MATH, kernel=KERNEL_RESIDENT, want_q=False)
print(o['stats'])
