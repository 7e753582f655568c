"""ctypes binding of the CPU oracle (oracle/build/libdecipher_oracle.so).  Test infrastructure:
imported only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs."""
import ctypes as C
import os
import subprocess

import numpy as np

from decipher_b200.capi import CatchmentDesc, NMEASURE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_libs = {}


def build_oracle():
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)


def load_oracle(opt="O2"):
    if opt in _libs:
        return _libs[opt]
    name = "libdecipher_oracle.so" if opt == "O2" else "libdecipher_oracle_O0.so"
    path = os.path.join(ORACLE_DIR, "build", name)
    if not os.path.exists(path):
        build_oracle()
    lib = C.CDLL(path)
    lib.dcp_oracle_run_ensemble.argtypes = [C.POINTER(CatchmentDesc), C.c_int64, C.c_void_p, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.dcp_oracle_ranmar.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p]
    lib.dcp_oracle_genpar.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.dcp_oracle_build_hist.argtypes = [C.POINTER(CatchmentDesc), C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                          C.POINTER(C.c_int32)]
    lib.dcp_oracle_route_timeseries.argtypes = [C.POINTER(CatchmentDesc), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p]
    _libs[opt] = lib
    return lib


def oracle_run(cat, mcpar, flags=0, want_q=True, want_perf=True, want_totals=True, n_threads=1, opt="O2"):
    """Runs the oracle on host arrays; mcpar (npt,7,M) F-ordered is modified in place (srinit clip)."""
    lib = load_oracle(opt)
    desc, keep = cat.to_desc()
    M = mcpar.shape[2]
    q = np.zeros((cat.nriv, cat.nstep, M), order="F") if want_q else None
    perf = np.zeros((NMEASURE, cat.nriv, M), order="F") if want_perf else None
    tot = np.zeros((3, cat.nriv, M), order="F") if want_totals else None
    diag = np.zeros((3, M), order="F")
    status = np.zeros(M, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    rc = lib.dcp_oracle_run_ensemble(C.byref(desc), M, p(mcpar), int(flags), p(q), p(perf), p(tot), p(diag),
                                     p(status), int(n_threads))
    assert rc == 0, rc
    del keep
    return dict(q=q, perf=perf, reach_totals=tot, init_diag=diag, status=status, mcpar=mcpar)


def oracle_ranmar(ij, kl, n_skip, n):
    out = np.zeros(n)
    load_oracle().dcp_oracle_ranmar(ij, kl, n_skip, n, out.ctypes.data_as(C.c_void_p))
    return out


def oracle_genpar(ij, kl, ll, ul, n_member):
    npt = ll.shape[0]
    ll = np.asfortranarray(ll, dtype=np.float64)
    ul = np.asfortranarray(ul, dtype=np.float64)
    out = np.zeros((npt, 7, n_member), order="F")
    load_oracle().dcp_oracle_genpar(ij, kl, npt, ll.ctypes.data_as(C.c_void_p), ul.ctypes.data_as(C.c_void_p),
                                    n_member, out.ctypes.data_as(C.c_void_p))
    return out


def oracle_build_hist(cat, v, T_cap=4096):
    desc, keep = cat.to_desc()
    hist = np.zeros((cat.nnode, T_cap))
    idx = np.zeros((cat.nnode, 2), dtype=np.int32)
    T = C.c_int32()
    rc = load_oracle().dcp_oracle_build_hist(C.byref(desc), float(v), T_cap, hist.ctypes.data_as(C.c_void_p),
                                             idx.ctypes.data_as(C.c_void_p), C.byref(T))
    del keep
    return rc, hist[:, :T.value].copy(), idx, T.value


def oracle_route_timeseries(cat, velocity, inflow, opt="O2"):
    """Standalone router of the oracle: velocity (n,), inflow (nriv, nstep, n) F-ordered -> routed, status."""
    desc, keep = cat.to_desc()
    velocity = np.ascontiguousarray(velocity, dtype=np.float64)
    n = velocity.shape[0]
    assert inflow.flags.f_contiguous and inflow.shape == (cat.nriv, cat.nstep, n)
    routed = np.zeros((cat.nriv, cat.nstep, n), order="F")
    status = np.zeros(n, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = load_oracle(opt).dcp_oracle_route_timeseries(C.byref(desc), n, p(velocity), p(inflow), p(routed), p(status))
    assert rc == 0, rc
    del keep
    return routed, status
