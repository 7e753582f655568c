import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from decipher_b200 import synth
from decipher_b200.capi import *
from oracle_binding import oracle_run
from parity import *
print("fp64 peak TF/s, ms:", measure_fp64_peak())
cat = synth.synth_catchment("C1")
mc = synth.sample_parameters(3, 10)
ref = oracle_run(cat, mc.copy(order='F'), flags=1)
dev = DeviceCatchment(cat)
out = dev.run(mc.copy(order='F'), flags=OPT_CHECK_BALANCE, want_q=True, want_totals=True)
print(out['stats'])
print("status", ref['status'], out['status'])
print("mcpar eq", np.array_equal(ref['mcpar'], out['mcpar']))
try:
    print(compare_flows(out['q'], ref['q']))
except AssertionError as e:
    print("FLOW MISMATCH", e)
print("perf ref", ref['perf'][:,:,0], "gpu", out['perf'][:,:,0])
print("tot ref", ref['reach_totals'][:,:,0], "gpu", out['reach_totals'][:,:,0])
print("diag ref", ref['init_diag'][:,:3], "gpu", out['init_diag'][:,:3])
# throughput probe
cat2 = synth.synth_catchment("C2", nstep=2048)
dev2 = DeviceCatchment(cat2)
for M in (4096, 32768, 131072):
    mc2 = synth.sample_parameters(1, M)
    for flags in (0, OPT_FAST_MATH):
        o = dev2.run(mc2.copy(order='F'), flags=flags, want_q=False)
        s = o['stats']
        print(M, "fast" if flags else "exact", "units/s (physics)", M*500*2048/(s['ms_physics']*1e-3), "total", M*500*2048/(s['ms_total']*1e-3), s)
