"""C4-like timing probe: nac HRUs, kW targets per HRU, M members, nstep steps through the C ABI (kernel AUTO).
usage: c4_probe.py nac kW M nstep [parity_members]"""
import os, sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from decipher_b200 import synth
from decipher_b200.capi import *
nac, kW, M, nstep = [int(x) for x in sys.argv[1:5]]
npar = int(sys.argv[5]) if len(sys.argv) > 5 else 0
t0 = time.time()
cat = synth.synth_catchment(nac=nac, nriv=5, npt=1, ngp=4, nge=1, nstep=nstep, kW=kW, seed=synth.BASE_SEED + 4, tree=True)
print("synth %.1f s, nnz(W) = %d (%.1f per HRU, density %.4f)" % (time.time() - t0, int(cat.ntrans.sum()), cat.ntrans.sum() / nac, cat.ntrans.sum() / nac / nac), flush=True)
dev = DeviceCatchment(cat)
KERNEL = int(os.environ.get('C4_KERNEL', '0'))
mc = synth.sample_parameters(1, M)
for flags, name in ((OPT_FAST_MATH, "fast"), (0, "exact")):
    for rep in range(2):
        o = dev.run(mc.copy(order='F'), flags=flags, want_q=False, kernel=KERNEL)
    s = o['stats']
    print("%s: %.3e units/s physics, %.3e total; per member-step %.1f us; %s" % (name, M * nac * nstep / (s['ms_physics'] * 1e-3), M * nac * nstep / (s['ms_total'] * 1e-3), s['ms_physics'] * 1e3 / nstep, s), flush=True)
if npar:
    from oracle_binding import oracle_run
    from parity import compare_flows, compare_perf
    mcp = synth.sample_parameters(1, npar)
    t0 = time.time()
    ref = oracle_run(cat, mcp.copy(order='F'), n_threads=16)
    print("oracle %.1f s" % (time.time() - t0))
    for flags, name in ((OPT_FAST_MATH, "fast"), (0, "exact")):
        out = dev.run(mcp.copy(order='F'), flags=flags, want_q=True, kernel=KERNEL)
        print(name, "parity", compare_flows(out['q'], ref['q']), compare_perf(out['perf'], ref['perf']), "kernel", out['stats']['kernel_used'])
dev.close()
