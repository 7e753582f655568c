"""Device accuracy of the FAST policy's branch-free primitives (dcp_fastmath.cuh) against NumPy/glibc."""
import ctypes as C

import numpy as np
import pytest

from decipher_b200.capi import load_library

pytestmark = pytest.mark.gpu


def _dev(op, x):
    lib = load_library()
    lib.dcp_debug_fastmath.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
    y = np.empty_like(x)
    rc = lib.dcp_debug_fastmath(op, x.size, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p))
    assert rc == 0, lib.dcp_last_error()
    return y


def _ulps(got, ref):
    return np.abs(got - ref) / np.spacing(np.abs(ref))


def test_fast_exp_log_rcp_within_two_ulp():
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-700, 700, 400000), rng.uniform(-1, 1, 200000), rng.uniform(-1e-6, 1e-6, 1000)])
    e = _ulps(_dev(0, x), np.exp(x))
    y = np.concatenate([np.exp(rng.uniform(-690, 690, 400000)), rng.uniform(0.5, 2.0, 200000)])
    l = np.abs(_dev(1, y) - np.log(y)) / np.maximum(np.spacing(np.abs(np.log(y))), 1e-300)
    r = _ulps(_dev(2, y), 1.0 / y)
    print("max ulp exp %.2f log %.2f rcp %.2f" % (e.max(), l[np.abs(np.log(y)) > 1e-3].max(), r.max()))
    assert e.max() <= 2.0
    assert l[np.abs(np.log(y)) > 1e-3].max() <= 2.0
    assert np.abs(_dev(1, y) - np.log(y))[np.abs(np.log(y)) <= 1e-3].max() < 1e-18
    assert r.max() <= 1.5


def test_table_driven_exp_log():
    """fm_exp_tab / fm_log_tab (the lean step of the second-generation resident kernel): 64- and 128-entry shared-memory
    tables + short polynomials.  exp within 2 ulp; log within 3 ulp where |log| > 1e-3 and 1e-18 absolute near 1."""
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-700, 700, 400000), rng.uniform(-1, 1, 200000), rng.uniform(-1e-6, 1e-6, 1000)])
    e = _ulps(_dev(3, x), np.exp(x))
    y = np.concatenate([np.exp(rng.uniform(-690, 690, 400000)), rng.uniform(0.5, 2.0, 200000), rng.uniform(0.999, 1.001, 50000)])
    ref = np.log(y)
    got = _dev(4, y)
    big = np.abs(ref) > 1e-3
    l = np.abs(got - ref)[big] / np.spacing(np.abs(ref[big]))
    print("table exp max ulp %.2f, table log max ulp %.2f, near 1 max abs %.2e" % (e.max(), l.max(), np.abs(got - ref)[~big].max()))
    assert e.max() <= 2.0 and l.max() <= 3.5
    assert np.abs(got - ref)[~big].max() < 1e-18
