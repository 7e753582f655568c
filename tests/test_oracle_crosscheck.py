"""The C++ oracle against the independent NumPy restatement (oracle/oracle_np.py) on small cases.
With no runnable reference and no upstream fixtures, two independently written restatements agreeing
to the last bits is the strongest pin available for the oracle itself."""
import os
import sys

import numpy as np
import pytest

from decipher_b200 import synth
from oracle_binding import oracle_build_hist, oracle_run

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import oracle_np  # noqa: E402


@pytest.mark.parametrize("kw", [dict(nac=12, nriv=2, npt=3, ngp=2, nge=1, nstep=400, kW=4, seed=7),
                                dict(nac=20, nriv=4, npt=2, ngp=1, nge=2, nstep=300, kW=5, seed=8, tree=True, ntt=2, dt=24.0)])
def test_cpp_oracle_matches_numpy_restatement(kw):
    cat = synth.synth_catchment(**kw)
    mc = synth.sample_parameters(cat.npt, 6)
    ref = oracle_run(cat, mc.copy(order="F"), flags=0, want_perf=False, want_totals=False)
    for m in range(6):
        q = oracle_np.run_member(cat, mc[:, :, m])
        a, b = q, ref["q"][:, :, m]
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin)
        # identical operations and libm: the only freedom is NumPy's vectorised exp/log vs scalar glibc (<= 1 ulp)
        scale = np.abs(b[fin]).max()
        assert np.abs(a[fin] - b[fin]).max() <= 1e-12 * scale, (m, np.abs(a[fin] - b[fin]).max(), scale)


def test_histogram_matches_numpy_restatement():
    cat = synth.synth_catchment(nac=10, nriv=4, npt=1, nstep=5, tree=True, seed=3)
    for v in (251.0, 999.5, 3999.0):
        rc, hist, idx, T = oracle_build_hist(cat, v)
        H, idx1, T2 = oracle_np.build_hist(cat, v)
        assert T == T2 and np.array_equal(idx[:, 0], idx1)
        assert np.allclose(hist, H, rtol=0, atol=1e-16)
