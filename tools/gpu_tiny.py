"""Tiny end-to-end runs of every kernel family (a few seconds in total), each checked against the oracle: the quickest
"did I break a kernel" probe on a GPU box (compute-sanitizer is not available on this pool).  tools/gpu_tiny.py [narrow wide stepwise extras]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from decipher_b200 import synth                                                    # noqa: E402
from decipher_b200.capi import (DeviceCatchment, KERNEL_RESIDENT, KERNEL_STEPWISE, KERNEL_STREAMED, OPT_CHECK_BALANCE, OPT_FAST_MATH,   # noqa: E402
                                select_behavioural)
from oracle_binding import oracle_run                                              # noqa: E402
from parity import compare_flows, compare_perf                                     # noqa: E402

which = sys.argv[1:] or ["narrow", "wide", "stepwise", "extras"]
if "narrow" in which:
    cat = synth.synth_catchment(nac=70, nriv=3, npt=2, ngp=2, nge=1, nstep=131, kW=4, seed=3, tree=True)
    mc = synth.sample_parameters(cat.npt, 11)
    ref = oracle_run(cat, mc.copy(order="F"))
    dev = DeviceCatchment(cat)
    for kernel, flags in ((KERNEL_STREAMED, OPT_CHECK_BALANCE), (KERNEL_RESIDENT, 0), (KERNEL_RESIDENT, OPT_FAST_MATH)):
        out = dev.run(mc.copy(order="F"), flags=flags, kernel=kernel, want_q=True, want_totals=True, want_sig=True)
        compare_flows(out["q"], ref["q"]); compare_perf(out["perf"], ref["perf"])
        print("narrow ok", kernel, flags, flush=True)
    dev.close()
if "wide" in which:
    cat = synth.synth_catchment(nac=530, nriv=2, npt=1, ngp=2, nge=1, nstep=70, kW=5, seed=5)
    mc = synth.sample_parameters(cat.npt, 5)
    ref = oracle_run(cat, mc.copy(order="F"))
    dev = DeviceCatchment(cat)
    out = dev.run(mc.copy(order="F"), flags=OPT_FAST_MATH, kernel=KERNEL_RESIDENT, want_q=True, want_sig=True)
    dev.close()
    compare_flows(out["q"], ref["q"]); compare_perf(out["perf"], ref["perf"])
    print("wide ok", flush=True)
if "stepwise" in which:
    os.environ["DCP_DENSE_MIN_DENSITY"] = "0.0"
    cat = synth.synth_catchment(nac=150, nriv=2, npt=1, ngp=2, nge=1, nstep=70, kW=20, seed=7)
    mc = synth.sample_parameters(cat.npt, 9)
    ref = oracle_run(cat, mc.copy(order="F"))
    dev = DeviceCatchment(cat)
    for flags in (0, OPT_FAST_MATH):
        out = dev.run(mc.copy(order="F"), flags=flags, kernel=KERNEL_STEPWISE, want_q=True, steps_per_tile=70)
        compare_flows(out["q"], ref["q"]); compare_perf(out["perf"], ref["perf"])
        print("stepwise ok", flags, flush=True)
    dev.close()
if "extras" in which:
    cat = synth.synth_catchment(nac=40, nriv=3, npt=1, ngp=2, nge=1, nstep=200, kW=4, seed=9, tree=True)
    mc = synth.sample_parameters(cat.npt, 1500)
    dev = DeviceCatchment(cat)
    perf = dev.run(mc, flags=OPT_FAST_MATH)["perf"]
    rng = np.random.default_rng(1)
    x = np.asfortranarray(rng.exponential(1e-4, (cat.nriv, cat.nstep, 5)))
    dev.route_timeseries(np.linspace(300.0, 3000.0, 5), x)
    dev.close()
    idx, w = select_behavioural(perf, 0, float(np.nanmedian(perf[0].min(axis=0))))
    assert 0 < len(idx) < 1500
    print("extras ok", flush=True)
print("TINY OK")
