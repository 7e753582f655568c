import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from decipher_b200 import synth
from decipher_b200.capi import *
from oracle_binding import oracle_run
from parity import *
cat = synth.synth_catchment("C2", nstep=1024)
mc = synth.sample_parameters(1, 100)
ref = oracle_run(cat, mc.copy(order='F'), n_threads=16)
dev = DeviceCatchment(cat)
for flags in (0, OPT_FAST_MATH):
    out = dev.run(mc.copy(order='F'), flags=flags, kernel=KERNEL_RESIDENT, want_q=True, want_totals=True)
    print(out['stats'])
    print("status eq", np.array_equal(ref['status'], out['status']))
    try:
        print(compare_flows(out['q'], ref['q']))
        print(compare_perf(out['perf'], ref['perf']))
    except AssertionError as e:
        print("MISMATCH", e)
cat2 = synth.synth_catchment("C2", nstep=4096)
dev2 = DeviceCatchment(cat2)
for M in (592, 2368):
    mc2 = synth.sample_parameters(1, M)
    for flags in (0, OPT_FAST_MATH):
        o = dev2.run(mc2.copy(order='F'), flags=flags, kernel=KERNEL_RESIDENT, want_q=False)
        s = o['stats']
        print(M, "fast" if flags else "exact", "units/s (physics)", M*500*4096/(s['ms_physics']*1e-3), "total", M*500*4096/(s['ms_total']*1e-3), s)
