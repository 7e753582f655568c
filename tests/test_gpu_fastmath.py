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
