"""ctypes binding of the C++ host library (csrc/host/dcp_host.cpp): the reference's file formats in,
the reference's output files out.  Mirrors the setup sequence of `program main`
(RRMODEL/dyna_main.f90:95-160) and its per-member output (dyna_write_output.f90)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .capi import Catchment, CatchmentDesc, DecipherError

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "csrc", "build", "libdecipher_host.so")
_lib = None


def load_host_library(path: str = HOST_LIB_PATH):
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise DecipherError(f"host library not built: {path} missing (run __graft_entry__.build())")
    lib = C.CDLL(path)
    lib.dcph_last_error.restype = C.c_char_p
    lib.dcph_load_project.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.dcph_read_auto_file.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.dcph_free.argtypes = [C.c_void_p]
    lib.dcph_desc.argtypes = [C.c_void_p]
    lib.dcph_desc.restype = C.POINTER(CatchmentDesc)
    lib.dcph_numsim.argtypes = [C.c_void_p]
    lib.dcph_seeds.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.dcph_bounds_ll.argtypes = [C.c_void_p]
    lib.dcph_bounds_ll.restype = C.POINTER(C.c_double)
    lib.dcph_bounds_ul.argtypes = [C.c_void_p]
    lib.dcph_bounds_ul.restype = C.POINTER(C.c_double)
    lib.dcph_out_prefix.argtypes = [C.c_void_p]
    lib.dcph_out_prefix.restype = C.c_char_p
    lib.dcph_genpar.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    lib.dcph_ranmar.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p]
    lib.dcph_write_member.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int]
    lib.dcph_parse_member_list.argtypes = [C.c_char_p, C.c_int64, C.c_void_p, C.c_int64]
    lib.dcph_parse_member_list.restype = C.c_int64
    lib.dcph_load_router.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.dcph_router_inflow.argtypes = [C.c_void_p]
    lib.dcph_router_inflow.restype = C.POINTER(C.c_double)
    lib.dcph_write_routed.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int, C.c_int]
    _lib = lib
    return lib


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


class Project:
    """A loaded DECIPHeR project (one catchment setup + one input setup)."""

    def __init__(self, workdir: str, id_project=1, id_cat=1, id_flow=1):
        self.lib = load_host_library()
        self._h = C.c_void_p()
        rc = self.lib.dcph_load_project(workdir.encode(), id_project, id_cat, id_flow, C.byref(self._h))
        if rc != 0:
            raise DecipherError(f"dcph_load_project: {self.lib.dcph_last_error().decode()}")
        d = self.lib.dcph_desc(self._h).contents
        self.desc = d
        self.numsim = self.lib.dcph_numsim(self._h)
        s1, s2 = C.c_int(), C.c_int()
        self.lib.dcph_seeds(self._h, C.byref(s1), C.byref(s2))
        self.seeds = (s1.value, s2.value)
        self.out_prefix = self.lib.dcph_out_prefix(self._h).decode()
        npt = d.npt
        self.ll = _arr(self.lib.dcph_bounds_ll(self._h), npt * 7, np.float64).reshape((npt, 7), order="F")
        self.ul = _arr(self.lib.dcph_bounds_ul(self._h), npt * 7, np.float64).reshape((npt, 7), order="F")

    def catchment(self) -> Catchment:
        d = self.desc
        nac, nriv, nnode = d.nac, d.nriv, d.nnode
        ntrans = _arr(d.ntrans, nac, np.int32)
        uc = _arr(d.uptree_count, nnode, np.int32)
        return Catchment(
            dt=d.dt, ntt=d.ntt, npt=d.npt, ac=_arr(d.ac, nac, np.float64), st=_arr(d.st, nac, np.float64),
            mslp=_arr(d.mslp, nac, np.float64), ippt=_arr(d.ippt, nac, np.int32), ipet=_arr(d.ipet, nac, np.int32),
            ipar=_arr(d.ipar, nac, np.int32), ipriv=_arr(d.ipriv, nac, np.int32), ims_rz=_arr(d.ims_rz, nac, np.int32),
            ims_uz=_arr(d.ims_uz, nac, np.int32), ntrans=ntrans, itrans=_arr(d.itrans, int(ntrans.sum()), np.int32),
            wtrans=_arr(d.wtrans, int(ntrans.sum()) + nac, np.float64),
            rivers=_arr(d.rivers, nac * nriv, np.float64).reshape((nac, nriv), order="F"),
            sum_ac_riv=_arr(d.sum_ac_riv, nriv, np.float64), qobs_start=_arr(d.qobs_start, nriv, np.float64),
            node_gauge_id=_arr(d.node_gauge_id, nnode, np.int32), node_type=_arr(d.node_type, nnode, np.int32),
            uptree_count=uc, uptree_index=_arr(d.uptree_index, int(uc.sum()), np.int32),
            frac_sub_area=_arr(d.frac_sub_area, nnode, np.float64),
            river_data=_arr(d.river_data, d.ncell * 8, np.float64).reshape((d.ncell, 8), order="F"),
            node_to_flow_mapping=_arr(d.node_to_flow_mapping, nriv, np.int32),
            rain=_arr(d.rain, d.ngp * d.nstep, np.float64).reshape((d.ngp, d.nstep), order="F"),
            pet=_arr(d.pet, d.nge * d.nstep, np.float64).reshape((d.nge, d.nstep), order="F"),
            qobs=_arr(d.qobs, nriv * d.nstep_obs, np.float64).reshape((nriv, d.nstep_obs), order="F"),
            flow_err=d.flow_err, t_warm=d.t_warm)

    def genpar(self, n_member=None) -> np.ndarray:
        n = self.numsim if n_member is None else int(n_member)
        out = np.zeros((self.desc.npt, 7, n), order="F")
        rc = self.lib.dcph_genpar(self._h, n, out.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise DecipherError(self.lib.dcph_last_error().decode())
        return out

    def write_member(self, i_mc, mcpar, q, totals, diag, status, f0_leading_zero=False):
        a = [np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel(order="F")) for x in (mcpar, q, totals, diag)]
        rc = self.lib.dcph_write_member(self._h, int(i_mc), *[x.ctypes.data_as(C.c_void_p) for x in a], int(status),
                                        int(bool(f0_leading_zero)))
        if rc != 0:
            raise DecipherError(self.lib.dcph_last_error().decode())

    def close(self):
        if self._h:
            self.lib.dcph_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RouterInputs:
    """The standalone router's inputs (DTA/route_run_standalone.f90) parsed by the C++ host: routing tree, river data and the
    ASCII-grid flow series; `desc` describes the river network for dcp_catchment_create / dcp_route_timeseries."""

    def __init__(self, river_data_file: str, flow_conn_file: str, timeseries_file: str):
        self.lib = load_host_library()
        self._h = C.c_void_p()
        rc = self.lib.dcph_load_router(river_data_file.encode(), flow_conn_file.encode(), timeseries_file.encode(), C.byref(self._h))
        if rc != 0:
            raise DecipherError(f"dcph_load_router: {self.lib.dcph_last_error().decode()}")
        self.desc = self.lib.dcph_desc(self._h).contents
        d = self.desc
        self.inflow = _arr(self.lib.dcph_router_inflow(self._h), d.nriv * d.nstep, np.float64).reshape((d.nriv, d.nstep), order="F")

    def network(self):
        d = self.desc
        uc = _arr(d.uptree_count, d.nnode, np.int32)
        return dict(nnode=d.nnode, nriv=d.nriv, nstep=d.nstep, node_gauge_id=_arr(d.node_gauge_id, d.nnode, np.int32),
                    node_type=_arr(d.node_type, d.nnode, np.int32), uptree_count=uc,
                    uptree_index=_arr(d.uptree_index, int(uc.sum()), np.int32),
                    node_to_flow_mapping=_arr(d.node_to_flow_mapping, d.nriv, np.int32),
                    river_data=_arr(d.river_data, d.ncell * 8, np.float64).reshape((d.ncell, 8), order="F"))

    def write_routed(self, routed: np.ndarray, path: str, precision: int = 6, f0_leading_zero: bool = False):
        a = np.ascontiguousarray(np.asarray(routed, dtype=np.float64).ravel(order="F"))
        if self.lib.dcph_write_routed(self._h, a.ctypes.data_as(C.c_void_p), path.encode(), precision, int(f0_leading_zero)) != 0:
            raise DecipherError(self.lib.dcph_last_error().decode())

    def close(self):
        if self._h:
            self.lib.dcph_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def parse_member_list(text: str, numsim: int) -> np.ndarray:
    """`decipher_b200_run -write-members` syntax -> ascending unique 0-based member indexes (raises on a bad list)."""
    lib = load_host_library()
    out = np.empty(max(int(numsim), 1), dtype=np.int64)
    n = lib.dcph_parse_member_list(text.encode(), int(numsim), out.ctypes.data_as(C.c_void_p), out.size)
    if n < 0:
        raise DecipherError(lib.dcph_last_error().decode())
    return out[:n].copy()


def ranmar(ij, kl, n_skip, n):
    out = np.zeros(n)
    load_host_library().dcph_ranmar(ij, kl, n_skip, n, out.ctypes.data_as(C.c_void_p))
    return out
