"""Flow-duration-curve signatures and GLUE-style behavioural filtering on the device (extensions, nothing
upstream).  The signatures come from a streaming histogram inside the kernels (no flow series leaves the chip); the NumPy
restatement (oracle/oracle_np.py) applied to the SAME flows must give the same numbers, applied to the oracle's flows nearly
the same (a flow within 1e-11 of a bin edge may change bins)."""
import os
import sys

import numpy as np
import pytest

from decipher_b200 import synth
from decipher_b200.capi import (DeviceCatchment, KERNEL_RESIDENT, KERNEL_STREAMED, KERNEL_STEPWISE, NMEASURE, OPT_FAST_MATH, PERF_KGE, PERF_NSE,
                                PERF_PBIAS, PERF_RMSE, select_behavioural)
from oracle_binding import oracle_run

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import oracle_np  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kernel,flags", [(KERNEL_RESIDENT, OPT_FAST_MATH), (KERNEL_RESIDENT, 0), (KERNEL_STREAMED, 0), (KERNEL_STEPWISE, OPT_FAST_MATH)])
def test_fdc_signatures_match_the_numpy_restatement(kernel, flags):
    cat = synth.synth_catchment("C2", nstep=3000)
    cat.t_warm = 200
    M = 45
    mc = synth.sample_parameters(cat.npt, M)
    dev = DeviceCatchment(cat)
    out = dev.run(mc.copy(order="F"), flags=flags, kernel=kernel, want_q=True, want_sig=True)
    only = dev.run(mc.copy(order="F"), flags=flags, kernel=kernel, want_q=False, want_sig=True)       # nothing of length nstep leaves the device
    dev.close()
    assert np.array_equal(out["sig"], only["sig"], equal_nan=True)
    want = oracle_np.fdc_signatures(out["q"], cat.t_warm)
    assert np.array_equal(np.isnan(want), np.isnan(out["sig"]))
    ok = ~np.isnan(want)
    assert np.abs(out["sig"][ok] - want[ok]).max() <= 1e-12 * np.abs(want[ok]).max()
    assert (out["sig"][0][ok[0]] >= out["sig"][1][ok[1]]).all() and (out["sig"][1][ok[1]] >= out["sig"][2][ok[2]]).all()   # Q5 >= Q50 >= Q95
    ref = oracle_run(cat, mc.copy(order="F"))
    want_ref = oracle_np.fdc_signatures(ref["q"], cat.t_warm)
    fin = ok & ~np.isnan(want_ref)
    rel = np.abs(out["sig"][:3][fin[:3]] - want_ref[:3][fin[:3]]) / np.maximum(np.abs(want_ref[:3][fin[:3]]), 1e-30)
    assert rel.max() < 1e-3, rel.max()


def test_behavioural_selection_matches_numpy():
    cat = synth.synth_catchment("C2", nstep=2500)
    M = 3000
    mc = synth.sample_parameters(cat.npt, M)
    dev = DeviceCatchment(cat)
    perf = dev.run(mc, flags=OPT_FAST_MATH)["perf"]
    dev.close()
    perf[0, 1, 7] = np.nan                                             # a NaN never passes
    for measure, gauge, hib in [(PERF_NSE, -1, True), (PERF_KGE, 2, True), (PERF_RMSE, 0, False), (PERF_PBIAS, -1, False)]:
        v = np.abs(perf[measure]) if measure == PERF_PBIAS else perf[measure]
        rows = v if gauge < 0 else v[gauge:gauge + 1]
        worst = np.nanmin(rows, axis=0) if hib else np.nanmax(rows, axis=0)
        thr = float(np.nanquantile(worst, 0.7 if hib else 0.3))         # about 30 % of the members are behavioural
        idx, w = select_behavioural(perf, measure, thr, gauge=gauge, higher_is_better=hib)
        with np.errstate(invalid="ignore"):
            ok = (rows >= thr).all(axis=0) if hib else (rows <= thr).all(axis=0)
        want = np.nonzero(ok)[0]
        assert np.array_equal(idx, want), (measure, len(idx), len(want))
        assert 0 < len(want) < M
        if measure == PERF_NSE:
            assert 7 not in want                                       # the NaN at gauge 1 fails the every-gauge rule
        dist = np.abs(rows - thr).min(axis=0)[want]
        ww = dist / np.add.accumulate(dist)[-1]
        assert np.allclose(w, ww, rtol=1e-13, atol=0) and abs(w.sum() - 1.0) < 1e-12
    idx, w = select_behavioural(perf, PERF_NSE, 2.0)                   # impossible threshold: empty selection
    assert len(idx) == 0
