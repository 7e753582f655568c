"""Independent NumPy restatement of the hot path, written from the reference sources a second time
(not translated from decipher_oracle.cpp) to cross-check the C++ oracle on small cases.

TEST INFRASTRUCTURE ONLY (same rule as decipher_oracle.cpp).  Vectorised over HRUs where the
reference's order of operations is element-wise, sequential (np.add.accumulate / Python loops) where
the reference accumulates.  Citations relative to /root/reference.
"""
from __future__ import annotations

import math

import numpy as np

F32 = lambda x: float(np.float32(x))          # default-real literals of the reference, widened


def seq_sum(a):
    a = np.asarray(a, dtype=np.float64)
    return float(np.add.accumulate(a)[-1]) if a.size else 0.0


def build_hist(cat, v):
    """route_processing_build_hist (DTA/dta_route_processing.f90:154-354) for velocity v [m per timestep]."""
    rd = cat.river_data
    nnode = cat.nnode
    min_outlet = rd[:, 6].min()
    rng_, reach_delay, tdd = [], np.zeros(nnode), np.zeros(nnode)
    for i in range(nnode):
        rows = np.where(np.rint(rd[:, 0]).astype(int) == cat.node_gauge_id[i])[0]
        s, e = rows[0], rows[0]
        while e + 1 < rd.shape[0] and int(round(rd[e + 1, 0])) == cat.node_gauge_id[i]:
            e += 1
        rng_.append((s, e))
        reach_delay[i] = (rd[s:e + 1, 2].max() - rd[s:e + 1, 6].min()) / v
        tdd[i] = (rd[s:e + 1, 6].min() - min_outlet) / v
    idx1 = np.floor(tdd).astype(int) + 1
    idx2 = np.ceil(reach_delay + tdd).astype(int) + 1
    T = int(math.ceil(max(1.0, (reach_delay + tdd).max()))) + 1
    H = np.zeros((nnode, T))
    for i in range(nnode):
        s, e = rng_[i]
        nb = idx2[i] - idx1[i] + 1
        dh = np.zeros(nb)
        for j in range(s, e + 1):
            a = rd[j, 3] / v
            b = (rd[j, 3] + (rd[j, 2] - rd[j, 6])) / v
            ha, hb = int(math.floor(a)) + 1, int(math.floor(b)) + 1
            if hb > nb:
                continue
            area = rd[j, 1]
            if ha == hb:
                dh[ha - 1] += area * (b - a)
            else:
                dh[ha:hb - 1] += area
                dh[ha - 1] += area * (1 - (a - ha + 1))
                dh[hb - 1] += area * (b - hb + 1)
        H[i, idx1[i] - 1:idx2[i]] = dh / seq_sum(dh)
    return H, idx1, T


def run_member(cat, mcpar):
    """One member: mainloop (RRMODEL/dyna_main_loop.f90:89-169).  mcpar (npt, 7).  Returns q (nriv, nstep)."""
    nac, nriv, nstep, ntt = cat.nac, cat.nriv, cat.nstep, cat.ntt
    dt = cat.dt
    ip = cat.ipar - 1
    szm, lnto, srmax, srinit, chv, td, smax = [mcpar[:, k].copy() for k in range(7)]
    t0dt = lnto + math.log(dt)                                           # dyna_param_setup.f90:66-67
    ac, st = cat.ac, cat.st
    H, idx1, T = build_hist(cat, chv[0] * dt)                            # dyna_init_satzone.f90:85-89
    F = np.zeros_like(H)
    # --- init_satzone (dyna_init_satzone.f90:92-213)
    mean_q = seq_sum(cat.qobs_start * cat.sum_ac_riv)
    meanatb, meanT0, meanSZM = seq_sum(st * ac), seq_sum(t0dt[ip] * ac), seq_sum(szm[ip] * ac)
    q0 = np.exp(t0dt[ip] - st)
    w_riv = np.array([cat.wtrans[o] for o in (np.cumsum(cat.ntrans + 1) - 1)])

    def qb_of(sd_, qbf_):
        tot = 0.0
        for ia in range(nac):
            for r in range(nriv):
                tot = tot + (dt * qbf_[ia] * w_riv[ia] * cat.rivers[ia, r] * ac[ia])
        return tot

    def calc_qb(D):
        sd_ = D + szm[ip] * ((meanatb - st) + (t0dt[ip] - meanT0))
        with np.errstate(over="ignore", invalid="ignore"):
            qbf_ = q0 * np.exp(-sd_ / szm[ip])
        neg, big = sd_ < 0, sd_ >= smax[ip]
        qbf_ = np.where(neg, q0, np.where(big, 0.0, qbf_))
        sd_ = np.where(neg, 0.0, np.where(big, smax[ip], sd_))
        return sd_, qbf_, qb_of(sd_, qbf_)

    q0bar = math.exp(meanT0 - meanatb)
    D = -meanSZM * math.log(mean_q / q0bar)
    guess, ndiff, qb_diff = mean_q, F32(0.0001), 1.0
    sd, qbf = np.zeros(nac), q0.copy()
    qb_tmp = qb_of(sd, qbf)
    if qb_tmp <= mean_q:
        qb_diff = 0.0
    it = 0
    while abs(qb_diff) > F32(0.01):
        it += 1
        if it > 100000000:
            break
        sd, qbf, qb_tmp = calc_qb(D)
        prev, qb_diff = qb_diff, (qb_tmp - mean_q) * 1000
        if qb_diff > F32(0.01):
            guess -= ndiff
        elif qb_diff < -F32(0.01):
            guess += ndiff
        D = -meanSZM * math.log(guess / q0bar) if guess > 0 else float("nan")
        if prev > 0 and qb_diff < 0:
            if ndiff > F32(1e-9):
                ndiff = ndiff / 10
            else:
                if not abs(qb_diff) < abs(prev):
                    sd, qbf, qb_tmp = calc_qb(D)
                break
    for l in np.unique(ip):
        if srinit[l] > srmax[l]:
            srinit[l] = srmax[l]
    qint1, suz, srz = qbf * ac, np.zeros(nac), srmax[ip] - srinit[ip]
    # --- Topmod (dyna_topmod.f90:86-233)
    dtt = dt / ntt
    cosm = np.cos(cat.mslp)
    q1 = q0 * cosm
    q2 = q0 * cosm * np.exp(-1 * cosm * smax[ip] / szm[ip])
    m2 = szm[ip] / cosm
    tr_off = np.concatenate(([0], np.cumsum(cat.ntrans)))
    w_off = np.concatenate(([0], np.cumsum(cat.ntrans + 1)))
    q = np.zeros((nriv, nstep))
    up_off = np.concatenate(([0], np.cumsum(cat.uptree_count)))
    with np.errstate(all="ignore"):
        for t in range(nstep):
            p, ep = cat.rain[cat.ippt - 1, t], cat.pet[cat.ipet - 1, t]
            qin = np.zeros(nac)                                           # dynamic_dist (dyna_dynamic_dist.f90:47-67)
            qb = np.zeros(nriv)
            for ia in range(nac):
                for k in range(cat.ntrans[ia]):
                    j = cat.itrans[tr_off[ia] + k] - 1
                    qin[j] = qin[j] + qbf[ia] * cat.wtrans[w_off[ia] + k] * ac[ia]
                for r in range(nriv):
                    qb[r] = qb[r] + dtt * qbf[ia] * w_riv[ia] * cat.rivers[ia, r] * ac[ia]
            # root_zone1 (dyna_root_zone1.f90:39-66)
            srz = srz + p
            pex = np.where(srz > srmax[ip], srz - srmax[ip], 0.0)
            srz = np.minimum(srz, srmax[ip])
            ea = ep * (srz / srmax[ip])
            ea = np.where(ea > ep, ep, ea)
            ea = np.where(ea >= srz, srz, ea)
            ea = np.where(ep > 0, ea, 0.0)
            srz = srz - ea
            # unsat_zone1 (dyna_unsat_zone1.f90:49-78)
            suz = suz + pex
            c1 = (sd <= 0) & (suz > 0)
            c2 = (~c1) & (sd > 0) & (suz > sd)
            exus = np.where(c1, suz, np.where(c2, suz - sd, 0.0))
            suz = np.where(c1, 0.0, np.where(c2, sd, suz))
            act = (sd > 0) & (suz > 0)
            uz2 = np.where(act, dt * (suz / (sd * td[ip])), 0.0)
            uz2 = np.where(act & (uz2 > suz), suz, uz2)
            suz = np.where(act, suz - uz2, suz)
            fl = act & (suz < 0.0000001)
            uz2 = np.where(fl, uz2 + suz, uz2)
            suz = np.where(fl, 0.0, suz)
            uz = np.where(uz2 > 0, uz2 / dtt, 0.0)
            # satzone_analyticalsolver (dyna_satzone_analyticalsolver.f90:49-119)
            sd0 = sd.copy()
            for itt in range(1, ntt + 1):
                qin2 = (qint1 + itt * (qin - qint1) / ntt) / ac
                u = qin2 + uz
                u2 = q2 + u
                a = u2 * dtt / m2
                e1 = 1 / np.exp(-sd / m2)
                small = m2 * np.log((e1 + q1 * dtt / m2) - a * (e1 - 0.5 * q1 * dtt / m2))
                reg = m2 * np.log(q1 / u2 + (e1 - q1 / u2) * np.exp(-a))
                zero = m2 * np.log(np.exp(sd / m2) + q1 * dtt / m2)
                inner = np.where(u2 > 0, np.where(a < F32(1e-10), small, reg), zero)
                tc = (sd - smax[ip]) / u
                e1s = 1 / np.exp(-smax[ip] / m2)
                late = m2 * np.log(q1 / u2 + (e1s - q1 / u2) * np.exp(-u2 * (dtt - tc) / m2))
                outer = np.where(tc > dtt, sd - u * dtt, late)
                sd = np.where(sd <= smax[ip], inner, outer)
                qbf = (sd - sd0) / dtt + uz + qin / ac
            exs = np.where(sd < 0, 0 - sd, 0.0)
            sd = np.where(sd < 0, 0.0, sd)
            over = qbf > q0
            exs = np.where(over, exs + (qbf - q0) * dtt, exs)
            qbf = np.where(over, q0, qbf)
            qint1 = qin.copy()
            ex = exs + exus                                               # results_ac (dyna_results_ac.f90:34-38)
            qof = np.zeros(nriv)
            for ia in range(nac):
                if ex[ia] > 0:
                    qof[cat.ipriv[ia] - 1] = qof[cat.ipriv[ia] - 1] + ex[ia] * ac[ia]
            qout = qb + qof                                               # river (dyna_river.f90:43)
            for i in range(nriv):                                         # route_process_run_step (dta_route_processing.f90:671-727)
                n = cat.node_to_flow_mapping[i] - 1
                F[n] = F[n] + H[n] * qout[i]
            for i in range(nriv):
                c = idx1[i] - 1
                s = 0.0 + F[i, c]
                for u_ in cat.uptree_index[up_off[i]:up_off[i + 1]]:
                    s = s + F[u_ - 1, c]
                q[i, t] = s / cat.frac_sub_area[cat.node_to_flow_mapping[i] - 1]
            F[:, 0] = 0
            F = np.roll(F, -1, axis=1)
    return q


# ------------------------------------------------------------------------------------------------
# Flow-duration-curve signatures (extension; include/decipher_b200.h DCP_SIG_*): NumPy restatement of the device's
# streaming histogram (bins cut on the IEEE bits: exponent + top three mantissa bits) and of its quantile rule.
# ------------------------------------------------------------------------------------------------
FDC_OCTAVE_LO, FDC_OCTAVES = -40, 48
FDC_BINS = 8 * FDC_OCTAVES + 2


def fdc_bins(q):
    q = np.ascontiguousarray(q, dtype=np.float64)
    key = (q.view(np.int64) >> 49).astype(np.int64) - (1023 + FDC_OCTAVE_LO) * 8
    b = np.where(key < 0, 0, np.where(key >= 8 * FDC_OCTAVES, FDC_BINS - 1, key + 1))
    return np.where(q > 0.0, b, 0)


def fdc_quantile(h, n_total, p):
    r = p * n_total
    cum = 0.0
    for b in range(FDC_BINS - 1, -1, -1):
        cnt = float(h[b])
        if cnt == 0.0:
            continue
        cum = cum + cnt
        if cum >= r:
            if b == 0:
                return 0.0
            key = b - 1
            lo = np.ldexp(1.0 + (key & 7) * 0.125, (key >> 3) + FDC_OCTAVE_LO)
            if b == FDC_BINS - 1:
                return float(np.ldexp(1.0, FDC_OCTAVE_LO + FDC_OCTAVES))
            hi = np.ldexp(1.0 + ((key & 7) + 1) * 0.125, (key >> 3) + FDC_OCTAVE_LO)
            return float(lo + ((cum - r) / cnt) * (hi - lo))
    return 0.0


def fdc_signatures(q, t_warm=0):
    """q: (nriv, nstep, M) -> (4, nriv, M): Q5, Q50, Q95 (flows exceeded 5 / 50 / 95 % of the time) and the mid-segment slope."""
    nriv, nstep, M = q.shape
    out = np.full((4, nriv, M), np.nan)
    for m in range(M):
        for g in range(nriv):
            series = q[g, t_warm:, m]
            if not np.isfinite(series).all():
                continue
            h = np.bincount(fdc_bins(series), minlength=FDC_BINS)
            n = float(h.sum())
            q33, q66 = fdc_quantile(h, n, 0.33), fdc_quantile(h, n, 0.66)
            out[0, g, m] = fdc_quantile(h, n, 0.05)
            out[1, g, m] = fdc_quantile(h, n, 0.50)
            out[2, g, m] = fdc_quantile(h, n, 0.95)
            out[3, g, m] = (np.log(q33) - np.log(q66)) / 0.33 if q66 > 0.0 else np.nan
    return out
