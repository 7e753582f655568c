"""Quick GPU check of the resident kernels: parity of the fast policy vs the oracle on C1/C2 windows, then throughput."""
import os, sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from decipher_b200 import synth
from decipher_b200.capi import *
from oracle_binding import oracle_run
from parity import *

def parity(name, nstep, M, **kw):
    cat = synth.synth_catchment(name, nstep=nstep) if name else synth.synth_catchment(nstep=nstep, **kw)
    mc = synth.sample_parameters(cat.npt, M)
    ref = oracle_run(cat, mc.copy(order='F'), n_threads=16)
    dev = DeviceCatchment(cat)
    for tot in (True, False):
        out = dev.run(mc.copy(order='F'), flags=OPT_FAST_MATH, kernel=KERNEL_RESIDENT, want_q=True, want_totals=tot)
        try:
            msg = (compare_flows(out['q'], ref['q']), compare_perf(out['perf'], ref['perf']))
            if tot:
                fin = np.isfinite(ref["q"]).all(axis=(0, 1))
                msg += (compare_close(out["reach_totals"][:, :, fin], ref["reach_totals"][:, :, fin], 1e-9, "totals"),)
            print(name or kw, "totals" if tot else "no-totals", "OK", msg, "status eq", np.array_equal(ref['status'], out['status']))
        except AssertionError as e:
            print(name or kw, "MISMATCH", e)
    dev.close()

if "--noparity" not in sys.argv:
    parity("C2", 1024, 100)
    parity("C1", 2000, 70)
    parity(None, 700, 40, nac=100, nriv=3, npt=2, ngp=2, nge=1, kW=6, seed=11, tree=True)
nstep = int(os.environ.get("QV_NSTEP", "4096"))
cat2 = synth.synth_catchment("C2", nstep=nstep)
dev2 = DeviceCatchment(cat2)
for M in (592, 2368):
    mc2 = synth.sample_parameters(1, M)
    for rep in range(2):
        o = dev2.run(mc2.copy(order='F'), flags=OPT_FAST_MATH, kernel=KERNEL_RESIDENT, want_q=False)
    s = o['stats']
    print(M, "fast units/s (physics) %.3e total %.3e" % (M*500*nstep/(s['ms_physics']*1e-3), M*500*nstep/(s['ms_total']*1e-3)), s)
