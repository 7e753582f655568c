"""The performance measures are an extension (nothing upstream computes them), so the oracle's moment-sum
implementation (oracle/decipher_oracle.cpp:objective) is pinned here against an independent NumPy evaluation of
the textbook definitions (np.corrcoef / std / mean) on the golden cases."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402
from decipher_b200.capi import NMEASURE, PERF_KGE, PERF_LOGNSE, PERF_NSE, PERF_PBIAS, PERF_RMSE  # noqa: E402
from oracle_binding import oracle_run  # noqa: E402


@pytest.mark.parametrize("name", list(make_golden.CASES))
def test_measures_match_textbook_definitions(name):
    cat, mc = make_golden.build(name)
    r = oracle_run(cat, mc.copy(order="F"))
    q, perf = r["q"], r["perf"]
    assert perf.shape == (NMEASURE, cat.nriv, mc.shape[2])
    n_t = min(cat.nstep, cat.qobs.shape[1])
    checked = 0
    for m in range(q.shape[2]):
        for g in range(cat.nriv):
            o_all = cat.qobs[g, cat.t_warm:n_t]
            s_all = q[g, cat.t_warm:n_t, m]
            valid = o_all != cat.flow_err
            if not np.isfinite(q[g, :, m]).all() or valid.sum() == 0:
                assert np.isnan(perf[:, g, m]).all()
                continue
            o, s = o_all[valid], s_all[valid]
            eps = o.mean() / 100
            lo, ls = np.log(np.maximum(o, eps)), np.log(np.maximum(s, eps))
            want = np.empty(NMEASURE)
            want[PERF_NSE] = 1 - ((s - o) ** 2).sum() / ((o - o.mean()) ** 2).sum()
            want[PERF_LOGNSE] = 1 - ((ls - lo) ** 2).sum() / ((lo - lo.mean()) ** 2).sum()
            rr = np.corrcoef(s, o)[0, 1]
            want[PERF_KGE] = 1 - np.sqrt((rr - 1) ** 2 + (s.std() / o.std() - 1) ** 2 + (s.mean() / o.mean() - 1) ** 2)
            want[PERF_PBIAS] = 100 * (s - o).sum() / o.sum()
            want[PERF_RMSE] = np.sqrt(((s - o) ** 2).mean())
            np.testing.assert_allclose(perf[:, g, m], want, rtol=1e-8, atol=1e-10)
            checked += 1
    assert checked > 0
