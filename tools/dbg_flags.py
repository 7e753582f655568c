import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from decipher_b200 import synth
from decipher_b200.capi import *
from oracle_binding import oracle_run
for ntt, mod in ((4, True), (1, True), (4, False)):
    cat = synth.synth_catchment(nac=48, nriv=2, npt=2, ngp=2, nge=1, nstep=400, kW=4, seed=21, ntt=ntt)
    if mod:
        cat.ims_uz = cat.ims_uz.copy(); cat.ims_uz[::5] = 2
        cat.ims_rz = cat.ims_rz.copy(); cat.ims_rz[3::7] = 0
    mc = synth.sample_parameters(cat.npt, 20)
    ref = oracle_run(cat, mc.copy(order="F"), flags=0)
    dev = DeviceCatchment(cat)
    for name, kernel, flags in (("str-exact", KERNEL_STREAMED, 0), ("res-exact", KERNEL_RESIDENT, 0), ("res-fast", KERNEL_RESIDENT, OPT_FAST_MATH), ("str-fast", KERNEL_STREAMED, OPT_FAST_MATH)):
        out = dev.run(mc.copy(order="F"), flags=flags, kernel=kernel, want_q=True)
        a, b = out["q"], ref["q"]
        scale = np.abs(b).mean(axis=1, keepdims=True)
        err = np.abs(a-b)/np.maximum(np.abs(b), 1e-3*scale+1e-300)
        pm = np.nanmax(err, axis=(0,1))
        w = np.unravel_index(np.nanargmax(err), err.shape)
        print(ntt, mod, name, "max", pm.max(), "member", pm.argmax(), "at", w, "q", b[w], "mean", scale[w[0],0,w[2]], "params", mc[:, :, pm.argmax()].ravel()[:14])
    dev.close()
