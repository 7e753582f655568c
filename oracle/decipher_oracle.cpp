// decipher_oracle.cpp — CPU restatement of DECIPHeR's Monte-Carlo rainfall-runoff hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
// library, and only as the checker or the timed CPU baseline.
//
// PARITY PIN: no Fortran compiler exists in this environment and the reference ships no fixtures for
// this path, so the reference cannot be built.  Its SOURCE is executed instead: oracle/f90_translate.py
// translates the reference's own .f90 files statement by statement, oracle/f90_runtime.py runs them
// with Fortran arithmetic semantics, tests/golden/make_f90_vectors.py commits the outputs
// (tests/golden/f90_exec/), and tests/test_f90_pin_cpu.py requires this file to reproduce them BIT FOR
// BIT (flows, reach totals, init block, parameter tables, flags, RANMAR stream, routing histograms).
// Not pinned: one compiler's code generation (FMA contraction, x87, -ffast-math).  Also pinned: RANMAR
// against its published known-answer test (James 1990; tests/test_oracle_cpu.py) and an independent
// NumPy restatement (oracle/oracle_np.py).
//
// Every routine follows the reference statement by statement: IEEE FP64, strict left-to-right
// evaluation, no FMA contraction (build with -ffp-contract=off), default-real literals taken as
// float32 widened to double, sums in ascending index order.  Citations are relative to
// /root/reference.
//
// Deliberate deviations (documented in DESIGN.md):
//   * `sumeout` (RRMODEL/dyna_results_ts.f90:28,68) and `tot_reach_ppt`
//     (RRMODEL/dyna_write_output.f90:55-57,63) are used uninitialised upstream; both start at 0 here.
//   * STOP conditions become per-member status bits; the member keeps running.
//   * the init search loop is capped at DCP_INIT_MAX_ITER = 1e8 iterations (hang guard; the executed
//     reference needs up to 2.7e6 on the test catchments).
//   * NSE / log-NSE do not exist upstream; they are defined in objective() below.

#include "../include/decipher_b200.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

// ----------------------------------------------------------------------------------------------
// float32 literals of the reference, widened exactly as gfortran does
// ----------------------------------------------------------------------------------------------
const double F32_1EM10  = (double)0.0000000001f;  // dyna_root_zone1.f90:74, dyna_unsat_zone1.f90:84, satzone :72
const double F32_1EM7   = (double)0.0000001f;     // dyna_satzone_analyticalsolver.f90:127
const double F32_1EM6   = (double)0.000001f;      // dyna_results_ts.f90:88,95
const double F32_1EM4   = (double)0.0001f;        // dyna_init_satzone.f90:116
const double F32_1EM2   = (double)0.01f;          // dyna_init_satzone.f90:134,142,146
const double F32_1EM9   = (double)1e-9f;          // dyna_init_satzone.f90:159

// dyna_hru_type (RRMODEL/dyna_common_types.f90:22-77), array-of-structs like the reference
struct Hru {
    double ac, st, mslp, q0;
    int ippt, ipet, ipar, ipriv, ims_rz, ims_uz, ntrans;
    const int32_t *itrans;   // 1-based targets
    const double *wtrans;    // ntrans+1 weights
    double pex, quz, p, ep, ea, uz, ex, exus, exs;
    double sd, suz, srz, qin, qint1, qbf, qbft1;
    double sum_pptin, sum_pein, sum_aeout, sumexs, sumexus, sdini, suzini, srzini, qbfini;
};

// dyna_riv_type (RRMODEL/dyna_common_types.f90:79-88)
struct Riv {
    double qb, qof, qout, sumqb, sumqof, sumqout;
};

// route_time_delay_hist_type (DTA/dta_route_processing.f90:33-48) + the per-node delays
struct Tdh {
    double v_param;                       // v_mode is always V_MODE_CONSTANT (dyna_read_routingdata.f90:92-93)
    int T = 0;                            // total_timesteps
    std::vector<double> hist;             // route_hist_table(nnode, T), row-major here: [node*T + col]
    std::vector<double> flow;             // flow_matrix(nnode, T)
    std::vector<int> idx1, idx2;          // node_hist_indexes(:,1), (:,2)   (1-based columns)
    std::vector<double> reach_delay, total_downstream_delay;
    bool dropped = false;
};

struct Ctx {
    const dcp_catchment_desc *d;
    std::vector<int> tr_off, w_off, up_off;   // CSR offsets of itrans, wtrans, uptree
};

inline double rd(const dcp_catchment_desc *d, int cell /*0-based*/, int col /*1-based*/) {
    return d->river_data[(size_t)(col - 1) * d->ncell + cell];
}

// DTA/dta_route_processing.f90:419-442 (V_MODE_CONSTANT branch)
inline double calculate_cell_delay(double dist, double v) { return dist / v; }

// DTA/dta_route_processing.f90:356-416
void hist_add_value(std::vector<double> &delay_hist, double cur_a, double cur_b, double value,
                    int bin_count, bool &dropped) {
    int hist_a = (int)std::floor(cur_a) + 1;
    int hist_b = (int)std::floor(cur_b) + 1;
    if (hist_b > bin_count) {          // :378-382  early return drops the cell
        dropped = true;
        return;
    }
    if (hist_a < 1) {                  // would be an out-of-bounds write upstream (:371 only stops on <0)
        dropped = true;
        return;
    }
    if (hist_a == hist_b) {            // :384-388
        double fraction_value = (cur_b - cur_a);
        delay_hist[hist_a - 1] = delay_hist[hist_a - 1] + value * fraction_value;
    } else {
        for (int k = hist_a + 1; k <= hist_b - 1; ++k)     // :405-407
            delay_hist[k - 1] = delay_hist[k - 1] + value;
        double fraction_value = (1 - (cur_a - hist_a + 1)); // :410
        delay_hist[hist_a - 1] = delay_hist[hist_a - 1] + (value * fraction_value);
        fraction_value = (cur_b - hist_b + 1);              // :413
        delay_hist[hist_b - 1] = delay_hist[hist_b - 1] + (value * fraction_value);
    }
}

// DTA/dta_route_processing.f90:154-354
void route_processing_build_hist(const Ctx &c, Tdh &tdh) {
    const dcp_catchment_desc *d = c.d;
    const int nnode = d->nnode, ncell = d->ncell;
    const double v = tdh.v_param;
    tdh.idx1.assign(nnode, 0);
    tdh.idx2.assign(nnode, 0);
    tdh.reach_delay.assign(nnode, 0.0);
    tdh.total_downstream_delay.assign(nnode, 0.0);
    std::vector<int> start_i(nnode, 0), end_i(nnode, 0);

    double min_dist_outlet = rd(d, 0, 7);                   // :184  minval(river_data(:,7))
    for (int j = 1; j < ncell; ++j) min_dist_outlet = std::min(min_dist_outlet, rd(d, j, 7));

    for (int i = 0; i < nnode; ++i) {                       // :186-243
        int riv_start_i = 0, riv_end_i = ncell;
        for (int j = 1; j <= ncell; ++j) {
            int id = (int)std::lround(rd(d, j - 1, 1));
            if (riv_start_i == 0) {
                if (id == d->node_gauge_id[i]) riv_start_i = j;
            } else if (id != d->node_gauge_id[i]) {
                riv_end_i = j - 1;
                break;
            }
        }
        if (riv_start_i == 0) continue;                     // :204-212 (fixtures never hit this)
        start_i[i] = riv_start_i;
        end_i[i] = riv_end_i;
        double max_dist = rd(d, riv_start_i - 1, 3), min_dist = rd(d, riv_start_i - 1, 7);
        for (int j = riv_start_i; j <= riv_end_i; ++j) {
            max_dist = std::max(max_dist, rd(d, j - 1, 3)); // :218
            min_dist = std::min(min_dist, rd(d, j - 1, 7)); // :220
        }
        double reach_dist = max_dist - min_dist;            // :226
        tdh.reach_delay[i] = calculate_cell_delay(reach_dist, v);          // :231
        min_dist = min_dist - min_dist_outlet;                              // :238
        tdh.total_downstream_delay[i] = calculate_cell_delay(min_dist, v); // :241
    }

    double headwater_max_delay = 1;                         // :247
    for (int i = 0; i < nnode; ++i) {                       // :253-270
        double full_delay = tdh.reach_delay[i] + tdh.total_downstream_delay[i];
        tdh.idx1[i] = (int)std::floor(tdh.total_downstream_delay[i]) + 1;
        tdh.idx2[i] = (int)std::ceil(full_delay) + 1;
        if (full_delay > headwater_max_delay) headwater_max_delay = full_delay;
    }
    const int T = (int)std::ceil(headwater_max_delay) + 1;  // :273
    tdh.T = T;
    tdh.hist.assign((size_t)nnode * T, 0.0);                // :285-292
    tdh.flow.assign((size_t)nnode * T, 0.0);

    std::vector<double> delay_hist;
    for (int i = 0; i < nnode; ++i) {                       // :294-352
        if (start_i[i] == 0) continue;
        const int hist_start_i = tdh.idx1[i], hist_end_i = tdh.idx2[i];
        const int bin_count = (hist_end_i - hist_start_i) + 1;
        delay_hist.assign(bin_count, 0.0);
        for (int j = start_i[i]; j <= end_i[i]; ++j) {
            double cell_dist = rd(d, j - 1, 3) - rd(d, j - 1, 7);   // :325
            double cell_dist_a = rd(d, j - 1, 4);                   // :333
            double cell_dist_b = cell_dist_a + cell_dist;           // :334
            double cur_a = calculate_cell_delay(cell_dist_a, v);
            double cur_b = calculate_cell_delay(cell_dist_b, v);
            hist_add_value(delay_hist, cur_a, cur_b, rd(d, j - 1, 2), bin_count, tdh.dropped);
        }
        double s = 0;                                       // :347  sum() then divide
        for (int k = 0; k < bin_count; ++k) s += delay_hist[k];
        for (int k = 0; k < bin_count; ++k) delay_hist[k] = delay_hist[k] / s;
        for (int k = 0; k < bin_count; ++k)                 // :349
            tdh.hist[(size_t)i * T + (hist_start_i - 1 + k)] = delay_hist[k];
    }
}

// DTA/dta_route_processing.f90:633-729
void route_process_run_step(const Ctx &c, Tdh &tdh, const double *step_flow_input,
                            double *step_flow_output) {
    const dcp_catchment_desc *d = c.d;
    const int T = tdh.T, nnode = d->nnode, nriv = d->nriv;
    for (int i = 0; i < nriv; ++i) {                        // :671-681
        int node_index = d->node_to_flow_mapping[i] - 1;
        double *f = &tdh.flow[(size_t)node_index * T];
        const double *h = &tdh.hist[(size_t)node_index * T];
        const double q = step_flow_input[i];
        for (int k = 0; k < T; ++k) f[k] = f[k] + h[k] * q;
    }
    for (int i = 0; i < nriv; ++i) {                        // :685-714  flow_output_indexes(i) = i
        int node_index = i;
        double q_sum = 0.0;
        int sum_index = tdh.idx1[node_index];
        q_sum = q_sum + tdh.flow[(size_t)node_index * T + (sum_index - 1)];
        for (int j = 0; j < d->uptree_count[node_index]; ++j) {
            int upstream_index = d->uptree_index[c.up_off[node_index] + j] - 1;
            q_sum = q_sum + tdh.flow[(size_t)upstream_index * T + (sum_index - 1)];
        }
        step_flow_output[i] = q_sum;
    }
    for (int n = 0; n < nnode; ++n) {                       // :723-727  zero column 1, cshift left
        double *f = &tdh.flow[(size_t)n * T];
        std::memmove(f, f + 1, sizeof(double) * (T - 1));
        f[T - 1] = 0.0;
    }
}

// RRMODEL/dyna_init_satzone.f90:298-347
double calculate_qbmax(const Ctx &c, std::vector<Hru> &h, double dt) {
    const dcp_catchment_desc *d = c.d;
    for (auto &x : h) { x.sd = 0; x.qbf = x.q0; }
    double qb_tmp = 0;
    for (int ia = 0; ia < d->nac; ++ia)
        for (int iriv = 0; iriv < d->nriv; ++iriv)
            qb_tmp = qb_tmp + (dt * h[ia].qbf * h[ia].wtrans[h[ia].ntrans]
                               * d->rivers[(size_t)iriv * d->nac + ia] * h[ia].ac);
    return qb_tmp;
}

// RRMODEL/dyna_init_satzone.f90:226-296
double calculate_qb(const Ctx &c, std::vector<Hru> &h, double D, const double *szm,
                    const double *t0dt, const double *smax, double meanatb, double meanT0DT,
                    double dt) {
    const dcp_catchment_desc *d = c.d;
    for (auto &x : h) {
        const int ip = x.ipar - 1;
        x.sd = D + szm[ip] * ((meanatb - x.st) + (t0dt[ip] - meanT0DT));        // :268
        x.qbf = x.q0 * std::exp(-x.sd / szm[ip]);                                // :269
        if (x.sd < 0) {
            x.sd = 0;
            x.qbf = std::exp(t0dt[ip] - x.st);
        } else if (x.sd >= smax[ip]) {
            x.sd = smax[ip];
            x.qbf = 0;
        }
    }
    double qb_tmp = 0;
    for (int ia = 0; ia < d->nac; ++ia)
        for (int iriv = 0; iriv < d->nriv; ++iriv)
            qb_tmp = qb_tmp + (dt * h[ia].qbf * h[ia].wtrans[h[ia].ntrans]
                               * d->rivers[(size_t)iriv * d->nac + ia] * h[ia].ac);
    return qb_tmp;
}

struct MemberOut {
    int status = 0;
    double diag[3] = {0, 0, 0};
};

// One ensemble member: the body of mainloop (RRMODEL/dyna_main_loop.f90:89-171).
void run_member(const Ctx &c, double *mcpar /*[npt x 7]*/, int flags, double *q /*[nriv x nstep]*/,
                double *reach_totals /*[3 x nriv] or null*/, MemberOut &out) {
    const dcp_catchment_desc *d = c.d;
    const int nac = d->nac, nriv = d->nriv, npt = d->npt, nstep = d->nstep, ntt = d->ntt;
    const double dt = d->dt;
    const bool check = (flags & DCP_OPT_CHECK_BALANCE) != 0;

    // ---- param_setup (RRMODEL/dyna_param_setup.f90:58-70) ----
    std::vector<double> szm(npt), lnto(npt), srmax(npt), srinit(npt), chv(npt), td(npt), smax(npt),
        chvdt(npt), t0dt(npt);
    for (int i = 0; i < npt; ++i) {
        szm[i] = mcpar[0 * npt + i];
        lnto[i] = mcpar[1 * npt + i];
        srmax[i] = mcpar[2 * npt + i];
        srinit[i] = mcpar[3 * npt + i];
        chv[i] = mcpar[4 * npt + i];
        td[i] = mcpar[5 * npt + i];
        smax[i] = mcpar[6 * npt + i];
        chvdt[i] = chv[i] * dt;
        t0dt[i] = lnto[i] + std::log(dt);
    }

    // ---- static HRU data + initialise_run (RRMODEL/dyna_initialise_run.f90:27-36) ----
    std::vector<Hru> h(nac);
    std::vector<Riv> riv(nriv);
    for (int ia = 0; ia < nac; ++ia) {
        Hru &x = h[ia];
        std::memset(&x, 0, sizeof(Hru));
        x.ac = d->ac[ia]; x.st = d->st[ia]; x.mslp = d->mslp[ia];
        x.ippt = d->ippt[ia]; x.ipet = d->ipet[ia]; x.ipar = d->ipar[ia]; x.ipriv = d->ipriv[ia];
        x.ims_rz = d->ims_rz[ia]; x.ims_uz = d->ims_uz[ia];
        x.ntrans = d->ntrans[ia];
        x.itrans = d->itrans + c.tr_off[ia];
        x.wtrans = d->wtrans + c.w_off[ia];
    }
    for (auto &r : riv) r = Riv{0, 0, 0, 0, 0, 0};

    // ---- init_satzone (RRMODEL/dyna_init_satzone.f90:85-222) ----
    Tdh tdh;
    tdh.v_param = chvdt[0];                                                  // :85
    route_processing_build_hist(c, tdh);                                     // :89
    if (tdh.dropped) out.status |= DCP_ST_HIST_DROPPED;

    double mean_q = 0;
    for (int iriv = 0; iriv < nriv; ++iriv)                                  // :92-95
        mean_q = mean_q + d->qobs_start[iriv] * d->sum_ac_riv[iriv];
    double meanatb = 0, meanT0DT = 0, meanSZM = 0;
    for (int ia = 0; ia < nac; ++ia) {                                       // :102-109
        const int ip = h[ia].ipar - 1;
        meanatb = meanatb + (h[ia].st * h[ia].ac);
        meanT0DT = meanT0DT + (t0dt[ip] * h[ia].ac);
        meanSZM = meanSZM + (szm[ip] * h[ia].ac);
        h[ia].q0 = std::exp(t0dt[ip] - h[ia].st);
    }
    const double q0 = std::exp(meanT0DT - meanatb);                          // :112
    double D = -meanSZM * std::log(mean_q / q0);                             // :113
    double guess = mean_q;
    double ndiff = F32_1EM4;
    double qb_diff = 1, qb_diff_prev;
    int sol_flag = 1;

    double qb_tmp = calculate_qbmax(c, h, dt);                               // :124
    if (qb_tmp <= mean_q) { qb_diff = 0; sol_flag = 2; }                     // :126-131

    long iter = 0;
    while (std::fabs(qb_diff) > F32_1EM2) {                                  // :134
        if (++iter > DCP_INIT_MAX_ITER) { out.status |= DCP_ST_INIT_ITERCAP; break; }
        qb_tmp = calculate_qb(c, h, D, szm.data(), t0dt.data(), smax.data(), meanatb, meanT0DT, dt);
        qb_diff_prev = qb_diff;
        qb_diff = (qb_tmp - mean_q) * 1000;                                  // :140
        if (qb_diff > F32_1EM2) guess = guess - ndiff;                       // :142-150
        else if (qb_diff < -F32_1EM2) guess = guess + ndiff;
        D = -meanSZM * std::log(guess / q0);                                 // :153
        if ((qb_diff_prev > 0) && (qb_diff < 0)) {                           // :156-172
            if (ndiff > F32_1EM9) {
                ndiff = ndiff / 10;
            } else {
                sol_flag = 0;
                if (std::fabs(qb_diff) < std::fabs(qb_diff_prev)) {
                    break;
                } else {
                    qb_tmp = calculate_qb(c, h, D, szm.data(), t0dt.data(), smax.data(), meanatb,
                                          meanT0DT, dt);
                    break;
                }
            }
        }
    }
    out.status |= (sol_flag & DCP_ST_SOLFLAG_MASK);
    out.diag[0] = mean_q; out.diag[1] = qb_tmp; out.diag[2] = guess;         // :188-190

    for (int ia = 0; ia < nac; ++ia) {                                       // :194-213
        Hru &x = h[ia];
        const int ip = x.ipar - 1;
        x.qbft1 = x.qbf;
        x.qint1 = x.qbft1 * x.ac;
        x.qbfini = x.qbf;
        x.suz = 0;
        if (srinit[ip] > srmax[ip]) {
            srinit[ip] = srmax[ip];
            mcpar[3 * npt + ip] = srinit[ip];
        }
        x.srz = (srmax[ip] - srinit[ip]);
        x.sdini = x.sd; x.srzini = x.srz; x.suzini = x.suz;
    }
    for (size_t i = 0; i < (size_t)nriv * nstep; ++i) q[i] = 0;              // :217-222

    // ---- Topmod (RRMODEL/dyna_topmod.f90:86-233) ----
    const double dtt = dt / (double)ntt;                                     // :86
    std::vector<double> qout(nriv);
    for (int it = 0; it < nstep; ++it) {
        for (auto &r : riv) { r.qof = 0; r.qb = 0; }                         // :96-97
        for (auto &x : h) x.qin = 0;                                         // :98
        for (auto &x : h) {                                                  // rain, dyna_rain.f90:27-30
            x.p = d->rain[(size_t)it * d->ngp + (x.ippt - 1)];
            x.ep = d->pet[(size_t)it * d->nge + (x.ipet - 1)];
        }
        // dynamic_dist (RRMODEL/dyna_dynamic_dist.f90:47-67)
        for (int ia = 0; ia < nac; ++ia) {
            Hru &x = h[ia];
            for (int iw = 0; iw < x.ntrans; ++iw) {
                Hru &t = h[x.itrans[iw] - 1];
                t.qin = t.qin + x.qbf * x.wtrans[iw] * x.ac;
            }
            for (int iriv = 0; iriv < nriv; ++iriv) {
                const double rv = d->rivers[(size_t)iriv * nac + ia];
                riv[iriv].qb = riv[iriv].qb + dtt * x.qbf * x.wtrans[x.ntrans] * rv * x.ac;
                riv[iriv].sumqb = riv[iriv].sumqb + dtt * x.qbf * x.wtrans[x.ntrans] * rv * x.ac;
            }
        }
        for (int ia = 0; ia < nac; ++ia) {                                   // dyna_topmod.f90:133-204
            Hru &x = h[ia];
            const int ip = x.ipar - 1;
            x.quz = 0; x.uz = 0; x.ex = 0; x.exs = 0; x.exus = 0;
            if (x.ims_rz == 1) {
                // root_zone1 (RRMODEL/dyna_root_zone1.f90:32-77)
                x.pex = 0;
                double storein = x.p, storeini = x.srz, storeout = 0;
                x.srz = x.srz + x.p;
                x.sum_pptin = x.sum_pptin + x.p;
                if (x.srz > srmax[ip]) {
                    x.pex = (x.srz - srmax[ip]);
                    x.srz = srmax[ip];
                }
                storeout = x.pex;
                x.ea = 0;
                if (x.ep > 0) {
                    x.ea = x.ep * (x.srz / srmax[ip]);
                    if (x.ea > x.ep) x.ea = x.ep;
                    if (x.ea >= x.srz) x.ea = x.srz;
                    x.srz = x.srz - x.ea;
                    x.sum_pein = x.sum_pein + x.ep;
                    x.sum_aeout = x.sum_aeout + x.ea;
                    storeout = storeout + x.ea;
                }
                double storeend = x.srz;
                double storebal = storeini + storein - storeout - storeend;
                // `- 0000000001` is the integer -1 (dyna_root_zone1.f90:74)
                if (check && (storebal > F32_1EM10 || storebal < -1.0)) out.status |= DCP_ST_RZ_BALANCE;
            }
            if (x.ims_uz == 1) {
                // unsat_zone1 (RRMODEL/dyna_unsat_zone1.f90:40-86)
                double uz2 = 0;
                x.uz = 0;
                x.suz = x.suz + x.pex;
                if (x.sd <= 0.0 && x.suz > 0.0) {
                    x.exus = x.suz;
                    x.sumexus = x.sumexus + (x.exus * x.ac);
                    x.suz = 0;
                } else if (x.sd > 0.0 && x.suz > x.sd) {
                    x.exus = (x.suz - x.sd);
                    x.sumexus = x.sumexus + (x.exus * x.ac);
                    x.suz = x.sd;
                }
                if (x.sd > 0.0 && x.suz > 0.0) {
                    uz2 = dt * (x.suz / (x.sd * td[ip]));
                    if (uz2 > x.suz) uz2 = x.suz;
                    x.suz = x.suz - uz2;
                    if (x.suz < 0.0000001) {
                        uz2 = uz2 + x.suz;
                        x.suz = 0;
                    }
                    x.quz = uz2;
                }
                if (uz2 > 0.0) x.uz = uz2 / dtt;
            }
            {
                // satzone_analyticalsolver (RRMODEL/dyna_satzone_analyticalsolver.f90:49-131)
                const double cosm = std::cos(x.mslp);
                const double q1 = x.q0 * cosm;
                const double q2 = x.q0 * cosm * std::exp(-1 * cosm * smax[ip] / szm[ip]);
                const double m2 = szm[ip] / cosm;
                const double storeini = x.sd;
                for (int itt = 1; itt <= ntt; ++itt) {
                    const double qin2 = (x.qint1 + itt * (x.qin - x.qint1) / ntt) / x.ac;   // :63
                    const double u = qin2 + x.uz;
                    const double u2 = q2 + u;
                    if (x.sd <= smax[ip]) {
                        if (u2 > 0) {
                            if (u2 * dtt / m2 < F32_1EM10) {                                  // :72-77
                                x.sd = m2 * std::log((1 / std::exp(-x.sd / m2) + q1 * dtt / m2)
                                                     - u2 * dtt / m2
                                                           * (1 / std::exp(-x.sd / m2) - 0.5 * q1 * dtt / m2));
                            } else {                                                          // :80
                                x.sd = m2 * std::log(q1 / u2 + (1 / std::exp(-x.sd / m2) - q1 / u2)
                                                                   * std::exp(-u2 * dtt / m2));
                            }
                        } else {                                                              // :84
                            x.sd = m2 * std::log(std::exp(x.sd / m2) + q1 * dtt / m2);
                        }
                    } else {                                                                  // :86-95
                        const double tc = (x.sd - smax[ip]) / u;
                        if (tc > dtt) {
                            x.sd = x.sd - u * dtt;
                        } else {
                            const double t_temp = dtt - tc;
                            const double S_temp = smax[ip];
                            x.sd = m2 * std::log(q1 / u2 + (1 / std::exp(-S_temp / m2) - q1 / u2)
                                                               * std::exp(-u2 * t_temp / m2));
                        }
                    }
                    x.qbf = (x.sd - storeini) / dtt + x.uz + x.qin / x.ac;                    // :98
                }
                if (x.sd < 0) {                                                               // :107-112
                    x.exs = 0 - x.sd;
                    x.sumexs = x.sumexs + (x.exs * x.ac);
                    x.sd = 0;
                }
                if (x.qbf > x.q0) {                                                           // :115-119
                    x.exs = x.exs + (x.qbf - x.q0) * dtt;
                    x.sumexs = x.sumexs + (((x.qbf - x.q0) * dtt) * x.ac);
                    x.qbf = x.q0;
                }
            }
            // results_ac (RRMODEL/dyna_results_ac.f90:27-38)
            x.qint1 = x.qin;
            x.qbft1 = x.qbf;
            x.ex = x.exs + x.exus;
            if (x.ex > 0.0) riv[x.ipriv - 1].qof = riv[x.ipriv - 1].qof + x.ex * x.ac;
        }
        // river (RRMODEL/dyna_river.f90:42-59)
        for (int iriv = 0; iriv < nriv; ++iriv) {
            riv[iriv].qout = riv[iriv].qb + riv[iriv].qof;
            riv[iriv].sumqout = riv[iriv].sumqout + riv[iriv].qout;
            riv[iriv].sumqof = riv[iriv].sumqof + riv[iriv].qof;
            qout[iriv] = riv[iriv].qout;
        }
        double *qt = q + (size_t)it * nriv;
        route_process_run_step(c, tdh, qout.data(), qt);
        for (int iriv = 0; iriv < nriv; ++iriv)
            qt[iriv] = qt[iriv] / d->frac_sub_area[d->node_to_flow_mapping[iriv] - 1];

        if (check) {
            // results_ts (RRMODEL/dyna_results_ts.f90:54-101), sumeout zero-initialised
            double sumqbf = 0, sumqbfini = 0, sumsd = 0, sumsrz = 0, sumsuz = 0;
            double sumsdini = 0, sumsrzini = 0, sumsuzini = 0, sumcatin = 0, sumeout = 0;
            double sumexs = 0, sumexus = 0, sumqb = 0, sumqof = 0, sumfout = 0;
            for (int ia = 0; ia < nac; ++ia) {
                const Hru &x = h[ia];
                sumqbf = sumqbf + x.qbf * x.ac;
                sumqbfini = sumqbfini + (x.qbfini * x.ac);
                sumsd = sumsd - (x.sd * x.ac);
                sumsrz = sumsrz + x.srz * x.ac;
                sumsuz = sumsuz + x.suz * x.ac;
                sumsrzini = sumsrzini + (x.srzini * x.ac);
                sumsuzini = sumsuzini + (x.suzini * x.ac);
                sumsdini = sumsdini - (x.sdini * x.ac);
                sumcatin = sumcatin + x.sum_pptin * x.ac;
                sumeout = sumeout + x.sum_aeout * x.ac;
                sumexs = sumexs + x.sumexs;
                sumexus = sumexus + x.sumexus;
            }
            const double sumallini = sumsrzini + sumsuzini + sumsdini;
            for (int ia = 0; ia < nriv; ++ia) {
                sumqb = sumqb + riv[ia].sumqb;
                sumqof = sumqof + riv[ia].sumqof;
                sumfout = sumfout + riv[ia].sumqout;
            }
            const double wbal = sumallini + sumcatin - sumeout - sumfout - (sumsrz + sumsuz + sumsd)
                                + ((sumqbfini - sumqbf) * dtt);
            if (std::fabs((sumexs + sumexus + sumqb) - sumfout) > F32_1EM6) out.status |= DCP_ST_TS_OUTSUM;
            if (std::fabs(wbal) > F32_1EM6) out.status |= DCP_ST_TS_WBAL;
            (void)sumqof;
        }
    }

    // ---- write_output reach totals (RRMODEL/dyna_write_output.f90:55-77), ppt zero-initialised ----
    if (reach_totals) {
        std::vector<double> ppt(nriv, 0.0), pet(nriv, 0.0), aet(nriv, 0.0), frac(nriv, 0.0);
        for (int ia = 0; ia < nac; ++ia) {
            const Hru &x = h[ia];
            const int r = x.ipriv - 1;
            ppt[r] = ppt[r] + (x.sum_pptin * x.ac);
            pet[r] = pet[r] + (x.sum_pein * x.ac);
            aet[r] = aet[r] + (x.sum_aeout * x.ac);
            frac[r] = frac[r] + x.ac;
        }
        for (int r = 0; r < nriv; ++r) {
            reach_totals[0 + 3 * r] = (ppt[r] / frac[r]) * 1000;   // [3 x nriv], measure fastest
            reach_totals[1 + 3 * r] = (pet[r] / frac[r]) * 1000;
            reach_totals[2 + 3 * r] = (aet[r] / frac[r]) * 1000;
        }
    }
    bool nonfinite = false;
    for (size_t i = 0; i < (size_t)nriv * nstep; ++i)
        if (!std::isfinite(q[i])) { nonfinite = true; break; }
    if (nonfinite) out.status |= DCP_ST_NONFINITE;
}

// Performance measures (EXTENSION — nothing upstream computes them).
// For gauge g: V = { t in (t_warm, min(nstep,nstep_obs)] : qobs(g,t) != flow_err }.
//   mean_o = sum_V qobs / |V|, eps = mean_o/100, sums ascending in t.
//   NSE    = 1 - sum_V (q-qobs)^2 / sum_V (qobs-mean_o)^2
//   logNSE = same on ln(max(x, eps)).
//   KGE    = 1 - sqrt((r-1)^2 + (alpha-1)^2 + (beta-1)^2) (Gupta et al. 2009) from the moment sums
//            sq = sum_V q, sqq = sum_V q*q, sqo = sum_V q*qobs:  mean_s = sq/n, var_s = sqq/n - mean_s^2,
//            var_o = sst/n, cov = sqo/n - mean_s*mean_o, r = cov/sqrt(var_s*var_o), alpha = sqrt(var_s/var_o),
//            beta = mean_s/mean_o
//   PBIAS  = 100*(sq - sum_V qobs)/sum_V qobs,   RMSE = sqrt(sse/n).
// |V| = 0, or any non-finite simulated flow at the gauge => NaN.
struct ObsStats {
    int n = 0;
    double mean = 0, sst = 0, eps = 0, lmean = 0, lsst = 0, sum = 0;
};

inline double lclip(double x, double eps) { return std::log(x > eps ? x : eps); }

ObsStats obs_stats(const dcp_catchment_desc *d, int g) {
    ObsStats s;
    const int n_t = std::min(d->nstep, d->nstep_obs);
    double sum = 0;
    for (int t = d->t_warm; t < n_t; ++t) {
        const double o = d->qobs[(size_t)t * d->nriv + g];
        if (o != d->flow_err) { sum = sum + o; s.n++; }
    }
    s.sum = sum;
    if (s.n == 0) return s;
    s.mean = sum / s.n;
    s.eps = s.mean / 100;
    double lsum = 0;
    for (int t = d->t_warm; t < n_t; ++t) {
        const double o = d->qobs[(size_t)t * d->nriv + g];
        if (o != d->flow_err) {
            s.sst = s.sst + (o - s.mean) * (o - s.mean);
            lsum = lsum + lclip(o, s.eps);
        }
    }
    s.lmean = lsum / s.n;
    for (int t = d->t_warm; t < n_t; ++t) {
        const double o = d->qobs[(size_t)t * d->nriv + g];
        if (o != d->flow_err) {
            const double lo = lclip(o, s.eps);
            s.lsst = s.lsst + (lo - s.lmean) * (lo - s.lmean);
        }
    }
    return s;
}

void objective(const dcp_catchment_desc *d, const std::vector<ObsStats> &os, const double *q,
               double *perf /*[DCP_NMEASURE x nriv]*/) {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    const int n_t = std::min(d->nstep, d->nstep_obs);
    for (int g = 0; g < d->nriv; ++g) {
        const ObsStats &s = os[g];
        double sse = 0, lsse = 0, sq = 0, sqq = 0, sqo = 0;
        bool bad = false;
        for (int t = 0; t < d->nstep; ++t)
            if (!std::isfinite(q[(size_t)t * d->nriv + g])) { bad = true; break; }
        for (int t = d->t_warm; t < n_t && !bad; ++t) {
            const double o = d->qobs[(size_t)t * d->nriv + g];
            if (o == d->flow_err) continue;
            const double x = q[(size_t)t * d->nriv + g];
            sse = sse + (x - o) * (x - o);
            const double dl = lclip(x, s.eps) - lclip(o, s.eps);
            lsse = lsse + dl * dl;
            sq = sq + x;
            sqq = sqq + x * x;
            sqo = sqo + x * o;
        }
        double *pf = perf + (size_t)DCP_NMEASURE * g;
        if (bad || s.n == 0) {
            for (int k = 0; k < DCP_NMEASURE; ++k) pf[k] = nan;
        } else {
            pf[DCP_PERF_NSE] = 1 - sse / s.sst;
            pf[DCP_PERF_LOGNSE] = 1 - lsse / s.lsst;
            const double nn = (double)s.n;
            const double mean_s = sq / nn;
            const double var_s = sqq / nn - mean_s * mean_s;
            const double var_o = s.sst / nn;
            const double cov = sqo / nn - mean_s * s.mean;
            const double r = cov / std::sqrt(var_s * var_o);
            const double alpha = std::sqrt(var_s / var_o);
            const double beta = mean_s / s.mean;
            pf[DCP_PERF_KGE] = 1.0 - std::sqrt(((r - 1.0) * (r - 1.0) + (alpha - 1.0) * (alpha - 1.0)) + (beta - 1.0) * (beta - 1.0));
            pf[DCP_PERF_PBIAS] = (100.0 * (sq - s.sum)) / s.sum;
            pf[DCP_PERF_RMSE] = std::sqrt(sse / nn);
        }
    }
}

Ctx make_ctx(const dcp_catchment_desc *d) {
    Ctx c;
    c.d = d;
    c.tr_off.resize(d->nac + 1);
    c.w_off.resize(d->nac + 1);
    c.tr_off[0] = 0;
    c.w_off[0] = 0;
    for (int ia = 0; ia < d->nac; ++ia) {
        c.tr_off[ia + 1] = c.tr_off[ia] + d->ntrans[ia];
        c.w_off[ia + 1] = c.w_off[ia] + d->ntrans[ia] + 1;
    }
    c.up_off.resize(d->nnode + 1);
    c.up_off[0] = 0;
    for (int n = 0; n < d->nnode; ++n) c.up_off[n + 1] = c.up_off[n] + d->uptree_count[n];
    return c;
}

// RANMAR (RRMODEL/dyna_random.f90:51-64, 111-133)
struct Ranmar {
    double cu[98];
    double cc, cd, cm;
    int i97, j97;
    void ranin(int ij, int kl) {
        int i = (ij / 177) % 177 + 2;
        int j = ij % 177 + 2;
        int k = (kl / 169) % 178 + 1;
        int l = kl % 169;
        for (int ii = 1; ii <= 97; ++ii) {
            double s = 0.0, t = 0.5;
            for (int jj = 1; jj <= 24; ++jj) {
                int m = (((i * j) % 179) * k) % 179;
                i = j; j = k; k = m;
                l = (53 * l + 1) % 169;
                if ((l * m) % 64 >= 32) s = s + t;
                t = 0.5 * t;
            }
            cu[ii] = s;
        }
        cc = 362436.0 / 16777216.0;
        cd = 7654321.0 / 16777216.0;
        cm = 16777213.0 / 16777216.0;
        i97 = 97;
        j97 = 33;
    }
    double next() {
        double uni = cu[i97] - cu[j97];
        if (uni < 0.0) uni = uni + 1.0;
        cu[i97] = uni;
        i97 = i97 - 1; if (i97 == 0) i97 = 97;
        j97 = j97 - 1; if (j97 == 0) j97 = 97;
        cc = cc - cd; if (cc < 0.0) cc = cc + cm;
        uni = uni - cc; if (uni < 0.0) uni = uni + 1.0;
        return uni;
    }
};

}  // namespace

extern "C" {

// n draws after skipping n_skip, from ranin(ij, kl)
void dcp_oracle_ranmar(int ij, int kl, int64_t n_skip, int64_t n, double *out) {
    Ranmar r;
    r.ranin(ij, kl);
    for (int64_t i = 0; i < n_skip; ++i) r.next();
    for (int64_t i = 0; i < n; ++i) out[i] = r.next();
}

// genpar for members 1..n_member (RRMODEL/dyna_genpar.f90:35-41): layer-outer, parameter-inner,
// one draw per cell; ll/ul/mcpar are [npt x 7] column-major per member.
void dcp_oracle_genpar(int ij, int kl, int npt, const double *ll, const double *ul,
                       int64_t n_member, double *mcpar) {
    Ranmar r;
    r.ranin(ij, kl);
    for (int64_t m = 0; m < n_member; ++m) {
        double *p = mcpar + (size_t)m * npt * 7;
        for (int ig = 0; ig < npt; ++ig)
            for (int igen = 0; igen < 7; ++igen) {
                const double u = r.next();
                p[igen * npt + ig] = ll[igen * npt + ig] + u * (ul[igen * npt + ig] - ll[igen * npt + ig]);
            }
    }
}

// Histogram of one velocity, for unit tests: hist [nnode x T_cap] row-major (node-major), idx [nnode x 2].
int dcp_oracle_build_hist(const dcp_catchment_desc *d, double v, int T_cap, double *hist, int32_t *idx,
                          int32_t *T_out) {
    Ctx c = make_ctx(d);
    Tdh tdh;
    tdh.v_param = v;
    route_processing_build_hist(c, tdh);
    *T_out = tdh.T;
    if (tdh.T > T_cap) return 1;
    for (int n = 0; n < d->nnode; ++n) {
        for (int k = 0; k < tdh.T; ++k) hist[(size_t)n * T_cap + k] = tdh.hist[(size_t)n * tdh.T + k];
        idx[2 * n] = tdh.idx1[n];
        idx[2 * n + 1] = tdh.idx2[n];
    }
    return tdh.dropped ? 2 : 0;
}

// Standalone router, same contract as dcp_route_timeseries (include/decipher_b200.h):
// route_processing_build_hist with v_param(1) = velocity[s], then route_process_run_timeseries
// (DTA/dta_route_processing.f90:583-630: zero the output, one route_process_run_step per row), as
// DTA/route_run_standalone.f90:118-189 drives them.  status: DCP_ST_HIST_DROPPED | DCP_ST_NONFINITE.
int dcp_oracle_route_timeseries(const dcp_catchment_desc *d, int64_t n_series, const double *velocity,
                                const double *inflow, double *routed, int32_t *status) {
    if (!d || d->struct_size != (int)sizeof(dcp_catchment_desc)) return DCP_ERR_INVALID;
    Ctx c = make_ctx(d);
    const size_t per = (size_t)d->nriv * d->nstep;
    for (int64_t s = 0; s < n_series; ++s) {
        Tdh tdh;
        tdh.v_param = velocity[s];
        route_processing_build_hist(c, tdh);
        int st = tdh.dropped ? DCP_ST_HIST_DROPPED : 0;
        for (int t = 0; t < d->nstep; ++t) {
            double *out = routed + s * per + (size_t)t * d->nriv;
            route_process_run_step(c, tdh, inflow + s * per + (size_t)t * d->nriv, out);
            for (int i = 0; i < d->nriv; ++i) if (!std::isfinite(out[i])) st |= DCP_ST_NONFINITE;
        }
        if (status) status[s] = st;
    }
    return DCP_OK;
}

// Same contract as dcp_run_ensemble (include/decipher_b200.h) on host buffers, n_threads worker
// threads over disjoint members (members are independent: SURVEY §8e).
int dcp_oracle_run_ensemble(const dcp_catchment_desc *d, int64_t n_member, double *mcpar, int flags,
                            double *q_out, double *perf_out, double *reach_totals, double *init_diag,
                            int32_t *status, int n_threads) {
    if (!d || d->struct_size != (int)sizeof(dcp_catchment_desc)) return DCP_ERR_INVALID;
    Ctx c = make_ctx(d);
    std::vector<ObsStats> os;
    if (perf_out && d->qobs) {
        for (int g = 0; g < d->nriv; ++g) os.push_back(obs_stats(d, g));
    }
    if (n_threads < 1) n_threads = 1;
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        std::vector<double> qloc;
        for (;;) {
            const int64_t m = next.fetch_add(1);
            if (m >= n_member) break;
            double *q = q_out ? q_out + (size_t)m * d->nriv * d->nstep : nullptr;
            if (!q) { qloc.resize((size_t)d->nriv * d->nstep); q = qloc.data(); }
            MemberOut mo;
            run_member(c, mcpar + (size_t)m * d->npt * 7, flags, q,
                       reach_totals ? reach_totals + (size_t)m * 3 * d->nriv : nullptr, mo);
            if (status) status[m] = mo.status;
            if (init_diag) for (int k = 0; k < 3; ++k) init_diag[(size_t)m * 3 + k] = mo.diag[k];
            if (perf_out) {
                double *pf = perf_out + (size_t)m * DCP_NMEASURE * d->nriv;
                if (d->qobs) objective(d, os, q, pf);
                else for (int k = 0; k < DCP_NMEASURE * d->nriv; ++k) pf[k] = std::numeric_limits<double>::quiet_NaN();
            }
        }
    };
    if (n_threads == 1) {
        worker();
    } else {
        std::vector<std::thread> th;
        for (int i = 0; i < n_threads; ++i) th.emplace_back(worker);
        for (auto &t : th) t.join();
    }
    return DCP_OK;
}

}  // extern "C"
