"""Runs shim/dyna_main_gpu.f90 -- the Fortran main program a maintainer would build against libdecipher_b200.so -- through the
Fortran-subset translator together with the reference's own modules, with the three C entry points it calls bound to Python
stand-ins: `dcp_catchment_create` rebuilds a Catchment from the bind(C) descriptor the Fortran code filled in (so the flattening
of dyna_hru / route_riv, the 1-based index conventions and the array orders are what is being tested), `dcp_run_ensemble` hands
it to a backend (the CPU oracle in the CPU test).  Test infrastructure; needs the reference tree."""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)

import f90_translate  # noqa: E402
import make_f90_vectors as mk  # noqa: E402
from decipher_b200 import capi  # noqa: E402


def load():
    prog = f90_translate.Program()
    for f in mk.FILES:
        prog.add_file(os.path.join(mk.REFERENCE, f))
    prog.add_file(os.path.join(ROOT, "shim", "decipher_c_api.f90"))
    prog.add_file(os.path.join(ROOT, "shim", "dyna_main_gpu.f90"))
    ns = prog.load()
    bad = [p for p in ("main", "write_output_from_arrays") if p in ns["_NOT_TRANSLATED"]]
    if bad:
        raise RuntimeError(f"shim procedures outside the subset: {bad} {ns['_NOT_TRANSLATED']}")
    return ns


def _np(x, dtype):
    return None if x is None or isinstance(x, int) else np.asarray(x.a, dtype=dtype)


def run_main(ns, root, backend, ids=(1, 1, 1)):
    """backend(cat, mcpar (npt,7,M)) -> dict(q, perf, reach_totals, init_diag, status, mcpar).
    -> dict(q (nriv,nstep,M) as written to unit 41, totals, flow_files, n_create, n_run)"""
    rt = ns["__runtime__"]
    rt.C_SIZEOF.update(T_dcp_catchment_desc=ctypes.sizeof(capi.CatchmentDesc), T_dcp_run_opts=ctypes.sizeof(capi.RunOpts))
    state = dict(cat=None, n_create=0, n_run=0)

    def create(d, handle):
        assert d.struct_size == ctypes.sizeof(capi.CatchmentDesc)
        f64, i32 = np.float64, np.int32
        nt = int(_np(d.ntrans, i32).sum())
        state["cat"] = capi.Catchment(
            dt=float(d.dt), ntt=int(d.ntt), npt=int(d.npt), ac=_np(d.ac, f64), st=_np(d.st, f64), mslp=_np(d.mslp, f64),
            ippt=_np(d.ippt, i32), ipet=_np(d.ipet, i32), ipar=_np(d.ipar, i32), ipriv=_np(d.ipriv, i32),
            ims_rz=_np(d.ims_rz, i32), ims_uz=_np(d.ims_uz, i32), ntrans=_np(d.ntrans, i32), itrans=_np(d.itrans, i32)[:nt],
            wtrans=_np(d.wtrans, f64), rivers=np.asfortranarray(_np(d.rivers, f64)), sum_ac_riv=_np(d.sum_ac_riv, f64),
            qobs_start=_np(d.qobs_start, f64), node_gauge_id=_np(d.node_gauge_id, i32), node_type=_np(d.node_type, i32),
            uptree_count=_np(d.uptree_count, i32), uptree_index=_np(d.uptree_index, i32)[:int(_np(d.uptree_count, i32).sum())],
            frac_sub_area=_np(d.frac_sub_area, f64), river_data=np.asfortranarray(_np(d.river_data, f64)),
            node_to_flow_mapping=_np(d.node_to_flow_mapping, i32), rain=np.asfortranarray(_np(d.rain, f64)),
            pet=np.asfortranarray(_np(d.pet, f64)), qobs=None, flow_err=float(d.flow_err), t_warm=int(d.t_warm))
        c = state["cat"]
        assert (c.nac, c.nriv, c.nstep) == (int(d.nac), int(d.nriv), int(d.nstep)) and c.nnode == int(d.nnode)
        assert c.river_data.shape == (int(d.ncell), 8) and c.rain.shape[0] == int(d.ngp) and c.pet.shape[0] == int(d.nge)
        state["n_create"] += 1
        return 0

    def run(handle, first, n, mcpar, opts, q_out, perf_out, totals, diag, status):
        assert opts.struct_size == ctypes.sizeof(capi.RunOpts) and int(first) == 1 and state["cat"] is not None
        assert mcpar.a.shape[2] == int(n)
        r = backend(state["cat"], np.asfortranarray(mcpar.a).copy(order="F"))
        mcpar.a[...] = r["mcpar"]                                   # in/out: the srinit clip
        for dst, key in ((q_out, "q"), (perf_out, "perf"), (totals, "reach_totals"), (diag, "init_diag"), (status, "status")):
            if isinstance(dst, rt.FArr):
                dst.a[...] = r[key]
        state["n_run"] += 1
        return 0

    ns.update(f_dcp_catchment_create=create, f_dcp_run_ensemble=run, f_dcp_catchment_destroy=lambda h: 0)
    auto = os.path.join(root, "auto_shim.dat")
    with open(auto, "w") as f:
        f.write("".join(f"{i}\n" for i in ids))
    rt.UNITS.clear()
    rt.IO.clear()
    rt.OUTPUT_FILES.clear()
    rt.ARGV[:] = ["decipher_gpu", "-auto", auto]
    cwd = os.getcwd()
    os.chdir(root)
    try:
        with np.errstate(all="ignore"):
            ns["f_main"]()
    except rt.FStop as e:
        if "STOP" not in str(e) or "failed" in str(e):
            raise
    finally:
        os.chdir(cwd)
    ev = [e for e in rt.IO.events if e[0].startswith("dyna_main_gpu.f90")]
    last = ev[-1][0]                                               # the unit-41 write of write_output_from_arrays
    rows = [e[2][0] for e in ev if e[0] == last]
    nriv = rows[0].shape[0]
    M = sum(1 for e in ev if e[1] == "! PARAMETERS")
    q = np.stack(rows, axis=1).reshape(nriv, M, -1).transpose(0, 2, 1)
    idx = [i for i, e in enumerate(ev) if e[1] in ("! PRECIP TOTALS (mm)", "! PET TOTALS (mm)", "! ET TOTALS (mm)")]
    tot = np.stack([ev[i + 1][2][0] for i in idx]).reshape(M, 3, nriv).transpose(1, 2, 0)
    return dict(q=q, totals=tot, flow_files=[f for u, f in rt.OUTPUT_FILES if u == 41], n_create=state["n_create"],
                n_run=state["n_run"], init=[e[2][0] for e in ev if e[1] == "Starting qobs "])
