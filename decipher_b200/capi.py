"""ctypes view of the C ABI in include/decipher_b200.h.

`Catchment` holds, as NumPy arrays in the reference's own (Fortran, 1-based-index) layouts, what
`program main` has in memory after its setup calls (RRMODEL/dyna_main.f90:95-160).  `load_library()`
loads the CUDA shared library; there is no CPU fallback — if the extension is missing or no B200 is
visible the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DCP_B200_LIB") or os.path.join(_HERE, "csrc", "build", "libdecipher_b200.so")   # override: kernel-tuning builds

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)

# error codes / flags (mirrors of the #defines)
DCP_OK = 0
DCP_ERR_INVALID, DCP_ERR_NO_DEVICE, DCP_ERR_CUDA, DCP_ERR_NOMEM, DCP_ERR_UNSUPPORTED = 1, 2, 3, 4, 5
ST_SOLFLAG_MASK = 0x3
ST_RZ_BALANCE = 0x4
ST_TS_OUTSUM = 0x8
ST_TS_WBAL = 0x10
ST_NONFINITE = 0x20
ST_INIT_ITERCAP = 0x40
ST_HIST_DROPPED = 0x80
OPT_CHECK_BALANCE = 0x1
OPT_DEVICE_BUFFERS = 0x2
OPT_FAST_MATH = 0x4
OPT_SIGNATURES = 0x8
NSIG = 4
SIG_Q5, SIG_Q50, SIG_Q95, SIG_SLOPE = range(4)
FDC_OCTAVE_LO, FDC_OCTAVES = -40, 48
FDC_BINS = 8 * FDC_OCTAVES + 2
KERNEL_AUTO, KERNEL_STREAMED, KERNEL_RESIDENT, KERNEL_STEPWISE = 0, 1, 2, 3
NMEASURE = 5
PERF_NSE, PERF_LOGNSE, PERF_KGE, PERF_PBIAS, PERF_RMSE = range(5)


class CatchmentDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("nac", C.c_int32), ("nriv", C.c_int32), ("npt", C.c_int32),
        ("nstep", C.c_int32), ("ngp", C.c_int32), ("nge", C.c_int32), ("ntt", C.c_int32),
        ("dt", C.c_double),
        ("ac", c_double_p), ("st", c_double_p), ("mslp", c_double_p),
        ("ippt", c_int32_p), ("ipet", c_int32_p), ("ipar", c_int32_p), ("ipriv", c_int32_p),
        ("ims_rz", c_int32_p), ("ims_uz", c_int32_p),
        ("ntrans", c_int32_p), ("itrans", c_int32_p), ("wtrans", c_double_p),
        ("rivers", c_double_p), ("sum_ac_riv", c_double_p), ("qobs_start", c_double_p),
        ("nnode", C.c_int32), ("ncell", C.c_int32),
        ("node_gauge_id", c_int32_p), ("node_type", c_int32_p),
        ("uptree_count", c_int32_p), ("uptree_index", c_int32_p),
        ("frac_sub_area", c_double_p), ("river_data", c_double_p),
        ("node_to_flow_mapping", c_int32_p),
        ("rain", c_double_p), ("pet", c_double_p),
        ("qobs", c_double_p), ("nstep_obs", C.c_int32), ("t_warm", C.c_int32),
        ("flow_err", C.c_double),
    ]


class RunOpts(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("flags", C.c_int32), ("kernel", C.c_int32),
                ("members_per_batch", C.c_int32), ("steps_per_tile", C.c_int32),
                ("reserved", C.c_int32), ("sig_out", C.c_void_p)]


class RunStats(C.Structure):
    _fields_ = [("ms_total", C.c_double), ("ms_init", C.c_double), ("ms_physics", C.c_double),
                ("ms_route", C.c_double), ("ms_h2d", C.c_double), ("ms_d2h", C.c_double),
                ("n_physics_launches", C.c_int64), ("n_launches", C.c_int64),
                ("kernel_used", C.c_int32), ("hist_len", C.c_int32)]


MAX_RANKS = 16


class MultiStats(C.Structure):
    _fields_ = [("n_ranks", C.c_int32), ("used_nccl", C.c_int32), ("kernel_used", C.c_int32), ("reserved", C.c_int32),
                ("ms_wall", C.c_double), ("ms_gather", C.c_double), ("ms_busy_max", C.c_double), ("ms_busy_sum", C.c_double),
                ("ms_rank_device", C.c_double * MAX_RANKS), ("ms_rank_wall", C.c_double * MAX_RANKS)]


def _f64(a, order="F"):
    return np.require(np.asarray(a, dtype=np.float64), requirements=["A", "O"] + (["F"] if order == "F" else ["C"]))


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


@dataclass
class Catchment:
    """In-memory catchment in the ABI's layouts (see dcp_catchment_desc for each field)."""
    dt: float
    ntt: int
    npt: int
    ac: np.ndarray
    st: np.ndarray
    mslp: np.ndarray
    ippt: np.ndarray
    ipet: np.ndarray
    ipar: np.ndarray
    ipriv: np.ndarray
    ims_rz: np.ndarray
    ims_uz: np.ndarray
    ntrans: np.ndarray
    itrans: np.ndarray
    wtrans: np.ndarray
    rivers: np.ndarray          # (nac, nriv)
    sum_ac_riv: np.ndarray
    qobs_start: np.ndarray
    node_gauge_id: np.ndarray
    node_type: np.ndarray
    uptree_count: np.ndarray
    uptree_index: np.ndarray
    frac_sub_area: np.ndarray
    river_data: np.ndarray      # (ncell, 8)
    node_to_flow_mapping: np.ndarray
    rain: np.ndarray            # (ngp, nstep)
    pet: np.ndarray             # (nge, nstep)
    qobs: Optional[np.ndarray] = None   # (nriv, nstep_obs)
    flow_err: float = -9999.0
    t_warm: int = 0
    meta: dict = field(default_factory=dict)

    @property
    def nac(self): return int(self.ac.shape[0])
    @property
    def nriv(self): return int(self.rivers.shape[1])
    @property
    def nstep(self): return int(self.rain.shape[1])
    @property
    def nnode(self): return int(self.node_gauge_id.shape[0])

    def to_desc(self):
        """Returns (desc, keepalive): keepalive owns the buffers the descriptor points into."""
        keep = {}

        def pd(name, a, order="F"):
            keep[name] = _f64(a, order)
            return keep[name].ctypes.data_as(c_double_p)

        def pi(name, a):
            keep[name] = _i32(a)
            return keep[name].ctypes.data_as(c_int32_p)

        d = CatchmentDesc()
        d.struct_size = C.sizeof(CatchmentDesc)
        d.nac, d.nriv, d.npt, d.nstep = self.nac, self.nriv, int(self.npt), self.nstep
        d.ngp, d.nge, d.ntt, d.dt = int(self.rain.shape[0]), int(self.pet.shape[0]), int(self.ntt), float(self.dt)
        d.ac, d.st, d.mslp = pd("ac", self.ac), pd("st", self.st), pd("mslp", self.mslp)
        d.ippt, d.ipet, d.ipar, d.ipriv = pi("ippt", self.ippt), pi("ipet", self.ipet), pi("ipar", self.ipar), pi("ipriv", self.ipriv)
        d.ims_rz, d.ims_uz = pi("ims_rz", self.ims_rz), pi("ims_uz", self.ims_uz)
        d.ntrans, d.itrans, d.wtrans = pi("ntrans", self.ntrans), pi("itrans", self.itrans), pd("wtrans", self.wtrans)
        d.rivers, d.sum_ac_riv, d.qobs_start = pd("rivers", self.rivers), pd("sum_ac_riv", self.sum_ac_riv), pd("qobs_start", self.qobs_start)
        d.nnode, d.ncell = self.nnode, int(self.river_data.shape[0])
        d.node_gauge_id, d.node_type = pi("node_gauge_id", self.node_gauge_id), pi("node_type", self.node_type)
        d.uptree_count, d.uptree_index = pi("uptree_count", self.uptree_count), pi("uptree_index", self.uptree_index)
        d.frac_sub_area, d.river_data = pd("frac_sub_area", self.frac_sub_area), pd("river_data", self.river_data)
        d.node_to_flow_mapping = pi("node_to_flow_mapping", self.node_to_flow_mapping)
        d.rain, d.pet = pd("rain", self.rain), pd("pet", self.pet)
        if self.qobs is not None:
            d.qobs = pd("qobs", self.qobs)
            d.nstep_obs = int(self.qobs.shape[1])
        else:
            d.qobs = None
            d.nstep_obs = 0
        d.t_warm = int(self.t_warm)
        d.flow_err = float(self.flow_err)
        return d, keep


class DecipherError(RuntimeError):
    pass


_lib = None


def load_library(path: str = LIB_PATH):
    """Loads libdecipher_b200.so (built by __graft_entry__.build()).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise DecipherError(
            f"CUDA extension not built: {path} missing. Run `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback.")
    lib = C.CDLL(path)
    lib.dcp_device_count.restype = C.c_int
    lib.dcp_set_device.argtypes = [C.c_int]
    lib.dcp_abi_version.restype = C.c_int
    lib.dcp_last_error.restype = C.c_char_p
    lib.dcp_catchment_create.argtypes = [C.POINTER(CatchmentDesc), C.POINTER(C.c_void_p)]
    lib.dcp_catchment_destroy.argtypes = [C.c_void_p]
    lib.dcp_run_ensemble.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.POINTER(RunOpts),
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dcp_get_run_stats.argtypes = [C.c_void_p, C.POINTER(RunStats)]
    lib.dcp_route_timeseries.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    lib.dcp_measure_fp64_peak.argtypes = [c_double_p, c_double_p]
    if hasattr(lib, "dcp_select_behavioural"):
        lib.dcp_select_behavioural.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32,
                                               C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int32]
    if not hasattr(lib, "dcp_comm_init"):        # a kernel-tuning build of an older ABI (DCP_B200_LIB): single-GPU entry points only
        _lib = lib
        return lib
    lib.dcp_comm_init.argtypes = [C.c_int, c_int32_p, C.POINTER(C.c_void_p)]
    lib.dcp_comm_destroy.argtypes = [C.c_void_p]
    lib.dcp_comm_size.argtypes = [C.c_void_p]
    lib.dcp_comm_uses_nccl.argtypes = [C.c_void_p]
    lib.dcp_run_ensemble_multi.argtypes = [C.c_void_p, C.POINTER(CatchmentDesc), C.c_int64, C.c_void_p, C.POINTER(RunOpts),
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(MultiStats)]
    _lib = lib
    return lib


def _check(lib, rc, what):
    if rc != DCP_OK:
        msg = lib.dcp_last_error()
        raise DecipherError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def make_opts(flags=0, kernel=KERNEL_AUTO, members_per_batch=0, steps_per_tile=0):
    o = RunOpts()
    o.struct_size = C.sizeof(RunOpts)
    o.flags, o.kernel = int(flags), int(kernel)
    o.members_per_batch, o.steps_per_tile = int(members_per_batch), int(steps_per_tile)
    return o


class DeviceCatchment:
    """Owns a dcp_catchment handle on the current CUDA device."""

    def __init__(self, cat: Catchment, device: Optional[int] = None):
        self.lib = load_library()
        if self.lib.dcp_device_count() < 1:
            raise DecipherError("no CUDA device visible; the DECIPHeR B200 path has no CPU fallback")
        if device is not None:
            _check(self.lib, self.lib.dcp_set_device(int(device)), "dcp_set_device")
        self.cat = cat
        desc, keep = cat.to_desc()
        self._h = C.c_void_p()
        _check(self.lib, self.lib.dcp_catchment_create(C.byref(desc), C.byref(self._h)), "dcp_catchment_create")
        del keep

    def close(self):
        if self._h:
            self.lib.dcp_catchment_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, mcpar: np.ndarray, *, flags=0, kernel=KERNEL_AUTO, want_q=False, want_perf=True,
            want_totals=False, want_sig=False, members_per_batch=0, steps_per_tile=0, first_member=1):
        """mcpar: (npt, 7, M) Fortran-ordered float64 host array (modified in place: srinit clip).
        Returns a dict of host arrays."""
        cat = self.cat
        assert mcpar.dtype == np.float64 and mcpar.flags.f_contiguous and mcpar.shape[:2] == (cat.npt, 7)
        M = mcpar.shape[2]
        out = {}
        q = np.empty((cat.nriv, cat.nstep, M), order="F") if want_q else None
        perf = np.empty((NMEASURE, cat.nriv, M), order="F") if want_perf else None
        tot = np.empty((3, cat.nriv, M), order="F") if want_totals else None
        diag = np.empty((3, M), order="F")
        status = np.zeros(M, dtype=np.int32)
        sig = np.empty((NSIG, cat.nriv, M), order="F") if want_sig else None
        opts = make_opts(flags | (OPT_SIGNATURES if want_sig else 0), kernel, members_per_batch, steps_per_tile)
        if want_sig:
            opts.sig_out = sig.ctypes.data_as(C.c_void_p)

        def p(a):
            return a.ctypes.data_as(C.c_void_p) if a is not None else None
        rc = self.lib.dcp_run_ensemble(self._h, int(first_member), int(M), p(mcpar), C.byref(opts),
                                       p(q), p(perf), p(tot), p(diag), p(status))
        _check(self.lib, rc, "dcp_run_ensemble")
        out.update(q=q, perf=perf, reach_totals=tot, init_diag=diag, status=status, mcpar=mcpar, sig=sig)
        out["stats"] = self.stats()
        return out

    def run_device(self, mcpar_ptr: int, M: int, *, perf_ptr=0, status_ptr=0, diag_ptr=0, totals_ptr=0,
                   q_ptr=0, flags=0, kernel=KERNEL_AUTO, members_per_batch=0, steps_per_tile=0):
        """Same call with DEVICE pointers (e.g. torch tensors' data_ptr()); nothing is copied."""
        opts = make_opts(flags | OPT_DEVICE_BUFFERS, kernel, members_per_batch, steps_per_tile)
        v = lambda x: C.c_void_p(x) if x else None
        rc = self.lib.dcp_run_ensemble(self._h, 1, int(M), v(mcpar_ptr), C.byref(opts), v(q_ptr), v(perf_ptr),
                                       v(totals_ptr), v(diag_ptr), v(status_ptr))
        _check(self.lib, rc, "dcp_run_ensemble(device)")
        return self.stats()

    def route_timeseries(self, velocity: np.ndarray, inflow: np.ndarray):
        """Standalone router (DTA/route_run_standalone.f90): velocity (n,) metres per timestep, inflow
        (nriv, nstep, n) Fortran-ordered.  Returns the routed series in the same layout, status words and timings."""
        cat = self.cat
        velocity = np.ascontiguousarray(velocity, dtype=np.float64)
        n = velocity.shape[0]
        assert inflow.dtype == np.float64 and inflow.flags.f_contiguous and inflow.shape == (cat.nriv, cat.nstep, n)
        routed = np.empty((cat.nriv, cat.nstep, n), order="F")
        status = np.zeros(n, dtype=np.int32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = self.lib.dcp_route_timeseries(self._h, n, p(velocity), p(inflow), p(routed), p(status), 0)
        _check(self.lib, rc, "dcp_route_timeseries")
        return dict(routed=routed, status=status, stats=self.stats())

    def route_timeseries_device(self, n: int, velocity_ptr: int, inflow_ptr: int, routed_ptr: int, status_ptr: int = 0):
        """Same call with DEVICE pointers."""
        v = lambda x: C.c_void_p(x) if x else None
        rc = self.lib.dcp_route_timeseries(self._h, int(n), v(velocity_ptr), v(inflow_ptr), v(routed_ptr), v(status_ptr),
                                           OPT_DEVICE_BUFFERS)
        _check(self.lib, rc, "dcp_route_timeseries(device)")
        return self.stats()

    def stats(self) -> dict:
        s = RunStats()
        _check(self.lib, self.lib.dcp_get_run_stats(self._h, C.byref(s)), "dcp_get_run_stats")
        return {k: getattr(s, k) for k, _ in RunStats._fields_}


class Communicator:
    """dcp_comm: the ranks (one device each, or several ranks on one device for testing) of a multi-GPU run in ONE process."""

    def __init__(self, n_ranks: int, devices=None):
        self.lib = load_library()
        self._h = C.c_void_p()
        dv = None
        if devices is not None:
            self._dev = _i32(devices)
            assert len(self._dev) == n_ranks
            dv = self._dev.ctypes.data_as(c_int32_p)
        _check(self.lib, self.lib.dcp_comm_init(int(n_ranks), dv, C.byref(self._h)), "dcp_comm_init")

    @property
    def uses_nccl(self):
        return bool(self.lib.dcp_comm_uses_nccl(self._h))

    def run(self, cat: Catchment, mcpar: np.ndarray, *, flags=0, kernel=KERNEL_AUTO, want_q=False, want_perf=True,
            want_totals=False, want_sig=False, members_per_batch=0):
        """dcp_run_ensemble_multi: the whole (npt, 7, M) table over every rank; same outputs as DeviceCatchment.run."""
        assert mcpar.dtype == np.float64 and mcpar.flags.f_contiguous and mcpar.shape[:2] == (cat.npt, 7)
        M = mcpar.shape[2]
        desc, keep = cat.to_desc()
        q = np.empty((cat.nriv, cat.nstep, M), order="F") if want_q else None
        perf = np.empty((NMEASURE, cat.nriv, M), order="F") if want_perf else None
        tot = np.empty((3, cat.nriv, M), order="F") if want_totals else None
        diag = np.empty((3, M), order="F")
        status = np.zeros(M, dtype=np.int32)
        sig = np.empty((NSIG, cat.nriv, M), order="F") if want_sig else None
        opts = make_opts(flags | (OPT_SIGNATURES if want_sig else 0), kernel, members_per_batch, 0)
        if want_sig:
            opts.sig_out = sig.ctypes.data_as(C.c_void_p)
        st = MultiStats()

        def p(a):
            return a.ctypes.data_as(C.c_void_p) if a is not None else None
        rc = self.lib.dcp_run_ensemble_multi(self._h, C.byref(desc), int(M), p(mcpar), C.byref(opts), p(q), p(perf), p(tot),
                                             p(diag), p(status), C.byref(st))
        del keep
        _check(self.lib, rc, "dcp_run_ensemble_multi")
        n = st.n_ranks
        stats = dict(n_ranks=n, used_nccl=bool(st.used_nccl), kernel_used=st.kernel_used, ms_wall=st.ms_wall, ms_gather=st.ms_gather,
                     ms_busy_max=st.ms_busy_max, ms_busy_sum=st.ms_busy_sum, ms_rank_device=list(st.ms_rank_device[:n]),
                     ms_rank_wall=list(st.ms_rank_wall[:n]))
        return dict(q=q, perf=perf, reach_totals=tot, init_diag=diag, status=status, mcpar=mcpar, sig=sig, stats=stats)

    def close(self):
        if self._h:
            self.lib.dcp_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def select_behavioural(perf: np.ndarray, measure: int, threshold: float, gauge: int = -1, higher_is_better: bool = True):
    """GLUE-style behavioural filter on the device (dcp_select_behavioural): perf (NMEASURE, nriv, M) Fortran-ordered host array.
    Returns (member indexes ascending, normalised likelihood weights)."""
    lib = load_library()
    assert perf.dtype == np.float64 and perf.flags.f_contiguous and perf.shape[0] == NMEASURE
    nriv, M = perf.shape[1], perf.shape[2]
    idx = np.empty(M, dtype=np.int64)
    wgt = np.empty(M, dtype=np.float64)
    n = C.c_int64()
    rc = lib.dcp_select_behavioural(perf.ctypes.data_as(C.c_void_p), M, nriv, int(measure), int(gauge), float(threshold),
                                    int(bool(higher_is_better)), idx.ctypes.data_as(C.c_void_p), wgt.ctypes.data_as(C.c_void_p),
                                    C.byref(n), 0)
    _check(lib, rc, "dcp_select_behavioural")
    return idx[:n.value].copy(), wgt[:n.value].copy()


def measure_fp64_peak():
    lib = load_library()
    tf, ms = C.c_double(), C.c_double()
    _check(lib, lib.dcp_measure_fp64_peak(C.byref(tf), C.byref(ms)), "dcp_measure_fp64_peak")
    return tf.value, ms.value
