"""Comparators for GPU-vs-oracle parity (tolerances of BASELINE.md §5)."""
import inspect
import json
import os

import numpy as np

RTOL = 1e-9      # north_star: flows and NSE/log-NSE within 1e-9 relative in FP64

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = os.path.join(_ROOT, "gpurun_out", "parity_report.jsonl")


def _caller():
    for fr in inspect.stack()[2:]:
        if fr.function.startswith("test_"):
            return f"{os.path.basename(fr.filename)}::{fr.function}"
    return "?"


def report(kind, **fields):
    """Achieved parity figures: printed (visible with `pytest -rP`) and appended to gpurun_out/parity_report.jsonl,
    which gpurun brings back from the GPU box (summaries are copied to profiles/)."""
    rec = {"test": _caller(), "kind": kind, **fields}
    print("[parity]", json.dumps(rec))
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        with open(REPORT, "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass
    return rec


def compare_flows(q_gpu, q_ref, rtol=RTOL, floor=1e-3):
    """q: (nriv, nstep, M).  Members whose oracle flows are non-finite are compared by class
    (finite / non-finite per member) only.  Returns a dict of diagnostics; raises AssertionError.
    Relative error is taken against max(|q_ref|, floor x mean|q_ref| of the gauge series): flows far
    below the series scale are differences of O(scale) terms (qb + qof, the store difference in
    dyna_satzone_analyticalsolver.f90:98) and carry the rounding noise of those terms."""
    assert q_gpu.shape == q_ref.shape
    M = q_ref.shape[2]
    fin_ref = np.isfinite(q_ref).all(axis=(0, 1))
    fin_gpu = np.isfinite(q_gpu).all(axis=(0, 1))
    assert np.array_equal(fin_ref, fin_gpu), f"finite-class mismatch: ref {np.where(~fin_ref)[0]}, gpu {np.where(~fin_gpu)[0]}"
    worst = 0.0
    worst_unfloored = 0.0        # |a-b|/|b| over every point with b != 0, no floor: reported, asserted at 1e3 x rtol
    worst_abs_floored = 0.0      # largest |a-b|/scale among the floored points
    n_floored = n_pts = n_zero_ref = 0
    for m in np.where(fin_ref)[0]:
        a, b = q_gpu[:, :, m], q_ref[:, :, m]
        scale = np.abs(b).mean(axis=1, keepdims=True)            # per gauge series scale
        thr = floor * scale + 1e-300
        d = np.abs(a - b)
        err = d / np.maximum(np.abs(b), thr)
        worst = max(worst, float(err.max()))
        fl = np.abs(b) < thr
        n_floored += int(fl.sum()); n_pts += b.size
        nz = b != 0.0
        n_zero_ref += int((~nz).sum())
        if nz.any():
            worst_unfloored = max(worst_unfloored, float((d[nz] / np.abs(b[nz])).max()))
        if fl.any():
            worst_abs_floored = max(worst_abs_floored, float((d / np.maximum(scale, 1e-300))[fl].max()))
        assert (a[~nz] == 0.0).all() or float(np.abs(a[~nz]).max()) <= rtol * float(scale.max()), "non-zero flow where the oracle has exact zeros"
    out = dict(max_rel=worst, max_rel_unfloored=worst_unfloored, n_floored=n_floored, n_points=n_pts, floor=floor,
               max_abs_over_scale_floored=worst_abs_floored, n_zero_ref=n_zero_ref, n_finite=int(fin_ref.sum()), n_members=M, rtol=rtol)
    report("flows", **out)
    assert worst <= rtol, f"max relative flow error {worst:.3e} > {rtol:.1e}"
    return out


def compare_perf(p_gpu, p_ref, rtol=RTOL):
    """p: (NMEASURE, nriv, M) = NSE, log-NSE, KGE, PBIAS, RMSE.  NaN / +-inf entries (no valid observations, zero
    variance of the observations, non-finite flows) must coincide by class.  Finite NSE / log-NSE / KGE are compared
    on 1 - measure (the accumulated ratio), RMSE relatively, PBIAS against the 100 % scale (it is a difference of two
    sums of flows, so its absolute error is rtol x 100 however small the bias is)."""
    assert p_gpu.shape == p_ref.shape
    nan_ref, nan_gpu = np.isnan(p_ref), np.isnan(p_gpu)
    assert np.array_equal(nan_ref, nan_gpu), "NaN pattern of performance measures differs"
    inf_ref = np.isinf(p_ref)
    assert np.array_equal(inf_ref, np.isinf(p_gpu)) and np.array_equal(np.sign(p_ref[inf_ref]), np.sign(p_gpu[inf_ref]))
    worst = 0.0
    for k in range(p_ref.shape[0]):
        ok = np.isfinite(p_ref[k])
        if not ok.any():
            continue
        a, b = p_gpu[k][ok], p_ref[k][ok]
        if k == 3:      # PBIAS
            err = np.abs(a - b) / np.maximum(np.abs(b), 100.0)
        elif k == 4:    # RMSE
            err = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
        else:
            err = np.abs(a - b) / np.maximum(np.abs(1.0 - b), 1e-300)
        assert err.max() <= rtol, f"measure {k}: max relative error {err.max():.3e} > {rtol:.1e}"
        worst = max(worst, float(err.max()))
    report("measures", max_rel=worst, n=int(np.isfinite(p_ref).sum()), rtol=rtol)
    return dict(max_rel=worst)


def compare_close(a, b, rtol=RTOL, what=""):
    a, b = np.asarray(a), np.asarray(b)
    both_bad = ~np.isfinite(a) & ~np.isfinite(b)
    err = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    err = np.where(both_bad, 0.0, err)
    assert np.nanmax(err) <= rtol and not np.isnan(err).any(), f"{what}: max rel err {np.nanmax(err):.3e}"
    return float(np.nanmax(err))
