"""Synthetic DECIPHeR catchments (the reference ships control files only, no data:
RRMODEL_EXAMPLE_FILES/settings.dat:11-15 points at files that are not in the repository).

Two layers:
  * `make_raw(...)`  — file-level values, already rounded to the precision the DTA writers print
    (DTA/dta_calculate_hrus.f90:327-354, 530-640; DTA/dta_preprocess.f90:279-341), so that the text
    files written by `decipher_b200.formats` parse back to exactly these numbers;
  * `assemble(raw)`  — applies the arithmetic the reference readers apply after parsing
    (RRMODEL/dyna_tread_dyna.f90:67-72,166-194; dyna_inputs.f90:84-113; dyna_read_routingdata.f90:52-89;
    DTA/dta_riv_tree_node.f90:80-110,563-590) and returns a `Catchment` in the C-ABI layouts.

Generator recipe: SURVEY.md §8(d).  Seeds: np.random.default_rng(20260119 + config_id).
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import numpy as np

from .capi import Catchment

BASE_SEED = 20260119

# RRMODEL_EXAMPLE_FILES/params.dat:7-9 — bounds per parameter layer, columns
# szm, lnto, srmax, srinit, chv, td, smax (order of dyna_param_setup.f90:58-64)
EXAMPLE_LL = np.array([[0.001, -7, 0.005, 0, 250, 0.1, 0.2],
                       [0.005, -7, 0.005, 0, 1000, 0.1, 0.2],
                       [0.01, 1, 0.005, 0, 3000, 0.1, 0.2]])
EXAMPLE_UL = np.array([[0.07, 5, 0.15, 0.01, 4000, 40, 3],
                       [0.007, -2, 0.15, 0.01, 4000, 1, 3],
                       [0.04, 4, 0.15, 0.01, 4000, 40, 3]])
EXAMPLE_SEEDS = (5, 2000)          # RRMODEL_EXAMPLE_FILES/params.dat:2


def example_bounds(npt: int):
    """(ll, ul) as (npt, 7) Fortran-ordered arrays; layers beyond 3 repeat layer 1."""
    rows = [i % 3 for i in range(npt)]
    ll = np.asfortranarray(EXAMPLE_LL[rows].astype(np.float64))
    ul = np.asfortranarray(EXAMPLE_UL[rows].astype(np.float64))
    return ll, ul


def seqsum(a) -> float:
    """Left-to-right IEEE sum (what a Fortran `sum()` / do-loop accumulation produces at -O0)."""
    a = np.asarray(a, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.add.accumulate(a)[-1])


def quant(x, dec):
    """Nearest double to the decimal with `dec` fractional digits — equals strtod('%.{dec}f' % x)."""
    return np.rint(np.asarray(x, dtype=np.float64) * 10.0 ** dec) / 10.0 ** dec


def make_raw(nac=24, nriv=2, npt=3, ngp=2, nge=1, nstep=8760, kW=4, seed=BASE_SEED, dt=1.0, ntt=1,
             reach_km=(5.0, 30.0), cell_m=50.0, tree=False, nstep_obs: Optional[int] = None,
             dense_w=False, qobs_fixture: Optional[str] = "auto") -> Dict:
    rng = np.random.default_rng(seed)
    raw: Dict = dict(nac=nac, nriv=nriv, npt=npt, ngp=ngp, nge=nge, nstep=nstep, dt=float(dt), ntt=int(ntt))

    # ---- HRU meta (HRU_ID NUM_CELLS MEAN_ATB MEAN_SLOPE GAUGE_ID RAIN_ID PET_ID PARAM_ID MODELSTRUCT_ID)
    st = np.sort(quant(rng.uniform(4.0, 12.0, nac), 5))             # ascending topographic index
    mslp = quant(rng.uniform(0.02, 0.6, nac), 8)
    cells = np.maximum(1, np.rint(rng.lognormal(5.0, 0.8, nac))).astype(np.int64)
    ipriv = rng.integers(1, nriv + 1, nac).astype(np.int32)
    ipriv[:nriv] = np.arange(1, nriv + 1)                             # every river drains at least one HRU
    ippt = rng.integers(1, ngp + 1, nac).astype(np.int32)
    ipet = rng.integers(1, nge + 1, nac).astype(np.int32)
    ipar = rng.integers(1, npt + 1, nac).astype(np.int32)
    ipar[:npt] = np.arange(1, npt + 1)                                # max(ipar) == npt (dyna_mc_setup.f90:89)
    rng.shuffle(ipriv); rng.shuffle(ipar)
    raw.update(st=st, mslp=mslp, cells=cells, ipriv=ipriv, ippt=ippt, ipet=ipet, ipar=ipar,
               ims=np.ones(nac, dtype=np.int32))

    # ---- HRU flux weights: self + (kW-1) down-slope targets (higher index = higher st), river weight
    ntrans = np.zeros(nac, dtype=np.int32)
    itrans, wtrans = [], []
    for i in range(nac):
        n_down = nac - 1 - i
        if dense_w:
            tg = np.arange(nac)                                      # config 4 "dense": every HRU
        else:
            k = min(kW - 1, n_down)
            tg = np.sort(rng.choice(np.arange(i + 1, nac), size=k, replace=False)) if k > 0 else np.array([], dtype=np.int64)
            tg = np.concatenate(([i], tg))
        w_riv = 0.01 + 0.29 * rng.uniform() * (0.3 + 0.7 * (st[i] - 4.0) / 8.0)
        w = rng.dirichlet(np.ones(len(tg))) * (1.0 - w_riv)
        w = quant(np.concatenate((w, [w_riv])), 6)
        if w[-1] <= 0:
            w[-1] = 1e-6
        ntrans[i] = len(tg)
        itrans.append((tg + 1).astype(np.int32))
        wtrans.append(w)
    raw.update(ntrans=ntrans, itrans=np.concatenate(itrans).astype(np.int32), wtrans=np.concatenate(wtrans))
    # each HRU drains to exactly one river with share 1
    raw["riv_share"] = [(int(ipriv[i]), 1.0) for i in range(nac)]

    # ---- river network: node 1 is the outlet; chain or random tree of nriv gauged reaches
    down = np.zeros(nriv, dtype=np.int32)                              # 1-based downstream node, 0 = none
    for n in range(1, nriv):
        down[n] = rng.integers(1, n + 1) if tree else n
    length = cell_m * np.rint(rng.uniform(reach_km[0], reach_km[1], nriv) * 1000.0 / cell_m)
    outlet_dist = np.zeros(nriv)
    for n in range(1, nriv):
        outlet_dist[n] = outlet_dist[down[n] - 1] + length[down[n] - 1]
    gauge_id = (1000 + np.arange(1, nriv + 1)).astype(np.int32)
    rows = []
    for n in range(nriv):
        nc = int(round(length[n] / cell_m))
        area = np.rint(rng.uniform(0.2, 3.0, nc) * cell_m * cell_m)
        elev0 = 10.0 + 0.002 * outlet_dist[n]
        for j in range(nc):
            ds_dist = outlet_dist[n] + j * cell_m
            dist = ds_dist + cell_m
            rows.append((gauge_id[n], area[j], dist, j * cell_m, elev0 + 0.002 * (j + 1) * cell_m, 0.002,
                         ds_dist, elev0 + 0.002 * j * cell_m))
    raw["river_data"] = np.asfortranarray(quant(np.array(rows, dtype=np.float64), 3))     # written with 3 decimals
    raw["flow_point"] = quant(np.array([(gauge_id[n], 1000.0 + n, 2000.0 + n, 2, outlet_dist[n], 10.0 + 0.002 * outlet_dist[n])
                                  for n in range(nriv)], dtype=np.float64), 3)
    raw["flow_conn"] = quant(np.array([(gauge_id[n], 1000.0 + n, 2000.0 + n, 2,
                                  gauge_id[down[n] - 1] if down[n] > 0 else 0,
                                  0.0, 0.0, length[down[n] - 1] if down[n] > 0 else 0.0, 1.0, 0.002)
                                 for n in range(nriv)], dtype=np.float64), 3)

    # ---- forcing: hourly rain (m/ts), PET, synthetic "observed" flow
    wet = rng.uniform(size=(ngp, nstep)) < 0.10
    rain = quant(np.where(wet, rng.exponential(1.2e-3, (ngp, nstep)), 0.0) * dt, 8)
    hours = np.arange(nstep) * dt
    pet1 = np.maximum(0.0, 1.5e-4 * np.sin(2 * np.pi * hours / 24.0 - np.pi / 2)) * (1 + 0.5 * np.sin(2 * np.pi * hours / (24.0 * 365)))
    pet = quant(np.stack([pet1 * (1.0 + 0.05 * g) for g in range(nge)]) * dt, 9)
    raw.update(rain=np.asfortranarray(rain), pet=np.asfortranarray(pet))

    n_obs = nstep if nstep_obs is None else nstep_obs
    # linear-reservoir proxy of net rainfall, 10 % multiplicative noise, 1 % missing
    net = np.maximum(rain.mean(axis=0) - 0.7 * pet.mean(axis=0), 0.0)
    qo = np.empty((nriv, n_obs))
    for g in range(nriv):
        k = 1.0 / (30.0 + 20.0 * g)
        s, base = 5e-5 / k, 2e-5 * dt
        series = np.empty(n_obs)
        for t in range(n_obs):
            s += net[t % nstep] * 0.6 - k * s
            series[t] = base + k * s
        qo[g] = series * (1 + 0.1 * rng.standard_normal(n_obs))
    qo = quant(np.maximum(qo, 1e-9), 9)
    flow_err = -9999.0
    qo[rng.uniform(size=qo.shape) < 0.01] = flow_err
    # BASELINE configs 2/3: the observed flow is an oracle run with mid-range parameters, kept as a data
    # fixture (decipher_b200/data/c2_qobs_oracle.npz, written by tests/golden/make_qobs_c2.py); it applies to exactly the
    # catchment it was generated for (seed, nstep, nriv), every other shape keeps the linear-reservoir proxy above
    if qobs_fixture == "auto":
        qobs_fixture = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "c2_qobs_oracle.npz")
    if qobs_fixture and os.path.exists(qobs_fixture) and nstep_obs is None:
        fx = np.load(qobs_fixture)
        if (int(fx["seed"]), int(fx["nstep"]), int(fx["nriv"])) == (int(seed), int(nstep), int(nriv)) and (nac, kW, dt) == (500, 8, 1.0):
            qe9 = fx["qobs_e9"].astype(np.float64)
            qo = np.where(qe9 < 0, flow_err, qe9 / 1e9)
            raw["qobs_source"] = "oracle run, mid-range parameters, 10 % noise, 1 % missing (data/c2_qobs_oracle.npz)"
    raw.update(qobs=np.asfortranarray(qo), flow_err=flow_err)
    return raw


def full_upstream(down: np.ndarray):
    """upstream_indexes + upstream_tree_indexes in the reference's construction order
    (build_upstream_from_downstream then riv_node_full_upstream,
    DTA/dta_riv_tree_node.f90:563-590 and :80-110).  `down` is 1-based, 0 = no downstream."""
    n = len(down)
    ups = [[] for _ in range(n)]
    for i in range(n):
        if down[i] > 0:
            ups[down[i] - 1].append(i + 1)
    tree = []

    def rec(acc, node):
        for u in ups[node - 1]:
            acc.append(u)
        for u in ups[node - 1]:
            rec(acc, u)
    for i in range(n):
        acc = []
        rec(acc, i + 1)
        tree.append(acc)
    return ups, tree


def assemble(raw: Dict) -> Catchment:
    nac, nriv = raw["nac"], raw["nriv"]
    cells = raw["cells"].astype(np.float64)
    # dyna_tread_dyna.f90:67-72
    tot = seqsum(cells)
    ac = cells / tot
    # HRU->river shares (dyna_tread_dyna.f90:157-170) and sum_ac_riv with the PRE-renormalisation ac
    rivers = np.zeros((nac, nriv), order="F")
    sum_ac_riv = np.zeros(nriv)
    for ia in range(nac):
        rid, share = raw["riv_share"][ia]
        rivers[ia, rid - 1] = share
        sum_ac_riv[raw["ipriv"][ia] - 1] += ac[ia]
    # weights renormalised if their sum is not exactly 1 (dyna_tread_dyna.f90:173-184)
    wtrans = raw["wtrans"].copy()
    off = 0
    for ia in range(nac):
        n = int(raw["ntrans"][ia]) + 1
        s = seqsum(wtrans[off:off + n])
        if s != 1.0:
            wtrans[off:off + n] = wtrans[off:off + n] / s
        off += n
    # ac renormalised inside a loop that recomputes sum(ac) every iteration (dyna_tread_dyna.f90:186-190)
    if seqsum(ac) != 1.0:
        for ia in range(nac):
            ac[ia] = ac[ia] / seqsum(ac)
    s_ac = seqsum(ac)
    sum_ac_riv = sum_ac_riv / s_ac                                        # :192-194

    # river tree (flow_point order = river index order)
    fp, fc = raw["flow_point"], raw["flow_conn"]
    gid = np.rint(fp[:, 0]).astype(np.int32)
    down = np.zeros(len(gid), dtype=np.int32)
    for row in fc:
        node = int(np.where(gid == int(round(row[0])))[0][0])
        hit = np.where(gid == int(round(row[4])))[0]
        down[node] = int(hit[0]) + 1 if len(hit) else 0
    _, tree = full_upstream(down)
    # frac_sub_area (dyna_read_routingdata.f90:52-89): own HRUs ascending, then each upstream node's
    frac = np.zeros(len(gid))
    ipriv = raw["ipriv"]
    for i in range(len(gid)):
        f = 0.0
        for z in np.where(ipriv == i + 1)[0]:
            f = ac[z] + f
        for u in tree[i]:
            for z in np.where(ipriv == u)[0]:
                f = ac[z] + f
        frac[i] = f

    # qobs_riv_step_start (dyna_inputs.f90:84-113)
    qobs, ferr, dt = raw["qobs"], raw["flow_err"], raw["dt"]
    start = qobs[:, 0].copy()
    for g in range(nriv):
        if start[g] == ferr:
            valid = qobs[g] != ferr
            if not valid.any():
                start[g] = float(np.float32(0.001) / np.float32(24)) * dt
            else:
                start[g] = seqsum(qobs[g][valid]) / int(valid.sum())

    return Catchment(
        dt=raw["dt"], ntt=raw["ntt"], npt=raw["npt"], ac=ac, st=raw["st"].copy(), mslp=raw["mslp"].copy(),
        ippt=raw["ippt"], ipet=raw["ipet"], ipar=raw["ipar"], ipriv=raw["ipriv"],
        ims_rz=np.ones(nac, dtype=np.int32), ims_uz=np.ones(nac, dtype=np.int32),
        ntrans=raw["ntrans"], itrans=raw["itrans"], wtrans=wtrans, rivers=rivers, sum_ac_riv=sum_ac_riv,
        qobs_start=start, node_gauge_id=gid, node_type=np.rint(fp[:, 3]).astype(np.int32),
        uptree_count=np.array([len(t) for t in tree], dtype=np.int32),
        uptree_index=np.array([u for t in tree for u in t], dtype=np.int32),
        frac_sub_area=frac, river_data=raw["river_data"],
        node_to_flow_mapping=np.arange(1, nriv + 1, dtype=np.int32),
        rain=raw["rain"], pet=raw["pet"], qobs=qobs, flow_err=ferr, t_warm=0,
        meta=dict(down=down, **({'qobs_source': raw['qobs_source']} if 'qobs_source' in raw else {})))


# Named configurations of BASELINE.json / SURVEY §8(d).  `members` is the ensemble size the config is
# quoted on; tests use smaller subsets.
CONFIGS = {
    "C1": dict(nac=24, nriv=2, npt=3, ngp=2, nge=1, nstep=8760, kW=4, seed=BASE_SEED + 1, members=10),
    "C2": dict(nac=500, nriv=3, npt=1, ngp=4, nge=1, nstep=87600, kW=8, seed=BASE_SEED + 2, members=100_000),
    "C3": dict(nac=500, nriv=3, npt=1, ngp=4, nge=1, nstep=87600, kW=8, seed=BASE_SEED + 2, members=1_000_000),
    "C4": dict(nac=5000, nriv=5, npt=1, ngp=4, nge=1, nstep=8760, kW=64, seed=BASE_SEED + 4, members=10_000, tree=True),
}


def synth_catchment(name: Optional[str] = None, **kw) -> Catchment:
    cfg = dict(CONFIGS[name]) if name else {}
    cfg.update(kw)
    cfg.pop("members", None)
    return assemble(make_raw(**cfg))


def sample_parameters(npt: int, n_member: int, seeds=EXAMPLE_SEEDS, ll=None, ul=None) -> np.ndarray:
    """RANMAR + genpar exactly as RRMODEL/dyna_random.f90:51-64,111-133 and dyna_genpar.f90:35-41:
    layer-outer / parameter-inner, one 24-bit draw per cell, sequential over members.
    Returns (npt, 7, n_member) Fortran-ordered.  (Product-side sampler used by bench.py and the
    Python host; the C++ host driver has the same generator in csrc/host.)"""
    if ll is None or ul is None:
        ll, ul = example_bounds(npt)
    ij, kl = seeds
    i = (ij // 177) % 177 + 2
    j = ij % 177 + 2
    k = (kl // 169) % 178 + 1
    l = kl % 169
    cu = np.zeros(98)
    for ii in range(1, 98):
        s, t = 0.0, 0.5
        for _ in range(24):
            m = (((i * j) % 179) * k) % 179
            i, j, k = j, k, m
            l = (53 * l + 1) % 169
            if (l * m) % 64 >= 32:
                s += t
            t *= 0.5
        cu[ii] = s
    # vectorising a lagged-Fibonacci generator is not worth it: work in 24-bit integers instead
    n = npt * 7 * n_member
    M24 = 1 << 24
    u = [int(round(x * M24)) for x in cu]
    cc, cd, cm = 362436, 7654321, 16777213
    i97, j97 = 97, 33
    out = np.empty(n, dtype=np.float64)
    for q in range(n):
        uni = u[i97] - u[j97]
        if uni < 0:
            uni += M24
        u[i97] = uni
        i97 -= 1
        if i97 == 0:
            i97 = 97
        j97 -= 1
        if j97 == 0:
            j97 = 97
        cc -= cd
        if cc < 0:
            cc += cm
        uni -= cc
        if uni < 0:
            uni += M24
        out[q] = uni / M24
    draws = out.reshape(n_member, npt, 7)                 # member, layer (outer), parameter (inner)
    ll = np.asarray(ll, dtype=np.float64)
    ul = np.asarray(ul, dtype=np.float64)
    mcpar = ll[:, :, None] + np.transpose(draws, (1, 2, 0)) * (ul - ll)[:, :, None]
    return np.asfortranarray(mcpar)
