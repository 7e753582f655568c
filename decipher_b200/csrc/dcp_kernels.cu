// dcp_kernels.cu — member-per-thread kernels of the ensemble path (sm_100a).
//
//   k_init_members      param_setup + init_satzone + route_processing_build_hist, one thread per member
//   k_physics_streamed  Topmod time loop over a time tile, state streamed through L2/HBM
//   k_route_conv        time-delay-histogram convolution of every node's own reach
//   k_route_gauge       upstream-tree summation, /frac_sub_area, flow output, NSE / log-NSE sums
//   k_finalize          reach totals + performance measures into the ABI's output layouts
//
// All per-member data are member-fastest, so every global access of a warp is one coalesced line and
// every catchment structure (W rows, river entries, forcing) is warp-uniform.
#include "dcp_device.cuh"
#include "dcp_physics.cuh"
#include "dcp_kernels.h"

#include <cfloat>
#include <cmath>

#define LDG(p) __ldg(p)

// ------------------------------------------------------------------------------------------------
// minimum over the batch of chv(layer 1): sizes the histogram buffers (slowest member = longest delay)
// ------------------------------------------------------------------------------------------------
__global__ void k_min_chv(const double *__restrict__ mcpar, int npt, int64_t n_member, double *out) {
    double v = DBL_MAX;
    for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < n_member;
         m += (int64_t)gridDim.x * blockDim.x) {
        const double x = mcpar[m * npt * 7 + 4 * npt + 0];
        v = (x < v) ? x : v;                     // NaN never selected
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_xor_sync(0xffffffffu, v, o);
        v = (w < v) ? w : v;
    }
    if ((threadIdx.x & 31) == 0) {
        // v > 0 for valid input: positive doubles order like their bit patterns
        if (v > 0.0) atomicMin((unsigned long long *)out, (unsigned long long)__double_as_longlong(v));
        else atomicMin((unsigned long long *)out, 0ull);
    }
}

// ------------------------------------------------------------------------------------------------
// K1: per-member initialisation — one WARP per member
// ------------------------------------------------------------------------------------------------
// The init search (dyna_init_satzone.f90:134-174) has a data-dependent trip count (a handful to
// thousands of iterations, each O(nac) with an exp per HRU), so a thread-per-member kernel is bound by
// its slowest thread.  Here the lanes of a warp share the HRUs of ONE member: calculate_qb runs 32 HRUs
// at a time and its river sum is a strided partial sum + shuffle tree (the only place where the
// summation order differs from the reference's ascending-HRU order; the search compares that sum with
// a 1e-5 tolerance, so a last-bit difference cannot change the path except at measure-zero ties).
// Everything order-sensitive that is cheap — the area-weighted means, the routing histogram — stays
// sequential on lane 0 in the reference's order.

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// calculate_qb (RRMODEL/dyna_init_satzone.f90:266-294): lanes stride over the HRUs
template <bool STORE>
__device__ __forceinline__ double calc_qb(const DevCatchment &C, const DevBatch &W, int m, int lane, double D,
                                          double meanatb, double meanT0DT) {
    const int B = W.B, npt = C.npt;
    double part = 0.0;
    for (int ia = lane; ia < C.nac; ia += 32) {
        const HruStatic h = C.hru[ia];
        const double szm = W.par[(0 * npt + h.ipar) * (size_t)B + m];
        const double smax = W.par[(6 * npt + h.ipar) * (size_t)B + m];
        const double t0 = W.t0dt[h.ipar * (size_t)B + m];
        const double q0 = W.q0[ia * (size_t)B + m];
        double sd = D + szm * ((meanatb - h.st) + (t0 - meanT0DT));           // :268
        double qbf = q0 * exp(-sd / szm);                                     // :269
        if (sd < 0.0) {
            sd = 0.0;
            qbf = q0;                 // exp(T0DT - st) again (:273): same expression as q0, same bits
        } else if (sd >= smax) {
            sd = smax;
            qbf = 0.0;
        }
        if (STORE) {
            W.sd[ia * (size_t)B + m] = sd;
            W.qbf0[ia * (size_t)B + m] = qbf;
        }
        const double base = C.dt * qbf * h.w_riv;                             // :289  (dt*qbf)*w
        for (int e = h.riv_beg; e < h.riv_end; ++e) {
            const RivEntry r = C.riv[e];
            part = part + (base * r.share * h.ac);
        }
    }
    return warp_sum(part);
}

// route_processing_build_hist + hist_add_value (DTA/dta_route_processing.f90:154-354, 356-416) for one member / one
// standalone series with v = metres travelled per timestep (calculate_cell_delay, :419-442, V_MODE_CONSTANT), in the
// reference's order; returns DCP_ST_* bits
__device__ __forceinline__ int build_hist_one(const DevCatchment &C, const DevBatch &W, const int m, const double chvdt1) {
    const int B = W.B;
    int status = 0;
    for (int n = 0; n < C.nnode; ++n) {
        const int cb = C.node_cell_beg[n], ce = C.node_cell_end[n];
        int hstart = 0, bins = 0;
        if (cb >= 0) {
            const double reach_delay = C.node_reach_dist[n] / chvdt1;     // :231
            const double tdd = C.node_down_dist[n] / chvdt1;              // :241
            const double full_delay = reach_delay + tdd;                  // :256
            hstart = (int)floor(tdd) + 1;                                 // :259
            const int hend = (int)ceil(full_delay) + 1;                   // :260
            bins = (hend - hstart) + 1;                                   // :312
            if (bins > W.Lcap || bins < 1) {                              // sized from the slowest member
                status |= DCP_ST_HIST_DROPPED;
                bins = bins < 1 ? 0 : W.Lcap;
            }
            double *dh = W.dh + ((size_t)m * C.nnode + n) * W.Lcap;       // [member][node][bin], bins contiguous
            for (int k = 0; k < bins; ++k) dh[k] = 0.0;
            for (int j = cb; j < ce; ++j) {
                const double cur_a = C.cell_a[j] / chvdt1;                // :337
                const double cur_b = C.cell_b[j] / chvdt1;                // :339
                const double value = C.cell_area[j];
                const int hist_a = (int)floor(cur_a) + 1;                 // hist_add_value :369-370
                const int hist_b = (int)floor(cur_b) + 1;
                if (hist_b > bins || hist_a < 1) {                        // :378-382 drops the cell
                    status |= DCP_ST_HIST_DROPPED;
                    continue;
                }
                if (hist_a == hist_b) {
                    const double fr = cur_b - cur_a;
                    dh[hist_a - 1] = dh[hist_a - 1] + value * fr;
                } else {
                    for (int k = hist_a + 1; k <= hist_b - 1; ++k) dh[k - 1] = dh[k - 1] + value;
                    double fr = (1.0 - (cur_a - (double)hist_a + 1.0));   // :410
                    dh[hist_a - 1] = dh[hist_a - 1] + (value * fr);
                    fr = (cur_b - (double)hist_b + 1.0);                  // :413
                    dh[hist_b - 1] = dh[hist_b - 1] + (value * fr);
                }
            }
            double s = 0.0;                                               // :347
            for (int k = 0; k < bins; ++k) s += dh[k];
            for (int k = 0; k < bins; ++k) dh[k] = dh[k] / s;
        }
        W.hstart[n * (size_t)B + m] = hstart;
        W.hbins[n * (size_t)B + m] = bins;
    }
    return status;
}

template <bool CHECK>
__global__ void __launch_bounds__(128)
k_init_members(DevCatchment C, DevBatch W, double *__restrict__ mcpar /*[npt x 7 x B]*/) {
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int B = W.B;
    if (m >= B) return;                                                       // whole warp leaves together
    const int npt = C.npt, nac = C.nac;
    int status = 0;

    // ---- param_setup (RRMODEL/dyna_param_setup.f90:58-70) ----
    double *mp = mcpar + (size_t)m * npt * 7;
    for (int i = lane; i < npt; i += 32) {
        for (int p = 0; p < 7; ++p) W.par[(p * npt + i) * (size_t)B + m] = mp[p * npt + i];
        W.t0dt[i * (size_t)B + m] = mp[1 * npt + i] + C.log_dt;
    }
    __syncwarp();
    const double chvdt1 = mp[4 * npt + 0] * C.dt;                             // dyna_init_satzone.f90:85

    // ---- route_processing_build_hist (DTA/dta_route_processing.f90:154-354), v = chvdt(1); lane 0, reference order ----
    if (lane == 0) status |= build_hist_one(C, W, m, chvdt1);

    // ---- init_satzone (RRMODEL/dyna_init_satzone.f90:92-213) ----
    // q0 (:108) and the hoisted q2 of dyna_satzone_analyticalsolver.f90:50-51 per HRU: independent, all lanes
    for (int ia = lane; ia < nac; ia += 32) {
        const HruStatic h = C.hru[ia];
        const double t0 = W.t0dt[h.ipar * (size_t)B + m];
        const double szm = W.par[(0 * npt + h.ipar) * (size_t)B + m];
        const double smax = W.par[(6 * npt + h.ipar) * (size_t)B + m];
        const double q0 = exp(t0 - h.st);
        W.q0[ia * (size_t)B + m] = q0;
        W.q2[ia * (size_t)B + m] = q0 * h.cosm * exp(-1.0 * h.cosm * smax / szm);
    }
    // the area-weighted means in ascending HRU order (:92-105) on lane 0, broadcast
    double mean_q = 0.0, meanatb = 0.0, meanT0DT = 0.0, meanSZM = 0.0;
    if (lane == 0) {
        for (int r = 0; r < C.nriv; ++r) mean_q = mean_q + C.qobs_start[r] * C.sum_ac_riv[r];
        for (int ia = 0; ia < nac; ++ia) {
            const HruStatic h = C.hru[ia];
            meanatb = meanatb + (h.st * h.ac);
            meanT0DT = meanT0DT + (W.t0dt[h.ipar * (size_t)B + m] * h.ac);
            meanSZM = meanSZM + (W.par[(0 * npt + h.ipar) * (size_t)B + m] * h.ac);
        }
    }
    mean_q = __shfl_sync(0xffffffffu, mean_q, 0);
    meanatb = __shfl_sync(0xffffffffu, meanatb, 0);
    meanT0DT = __shfl_sync(0xffffffffu, meanT0DT, 0);
    meanSZM = __shfl_sync(0xffffffffu, meanSZM, 0);
    __syncwarp();                                                             // q0 stores visible to every lane
    const double q0bar = exp(meanT0DT - meanatb);                             // :112
    double D = -meanSZM * log(mean_q / q0bar);                                // :113
    double guess = mean_q, ndiff = DCP_F32_1EM4, qb_diff = 1.0, qb_diff_prev;
    int sol_flag = 1;

    // calculate_qbmax (:324-345): sd = 0, qbf = q0
    double part = 0.0;
    for (int ia = lane; ia < nac; ia += 32) {
        const HruStatic h = C.hru[ia];
        const double base = C.dt * W.q0[ia * (size_t)B + m] * h.w_riv;
        for (int e = h.riv_beg; e < h.riv_end; ++e) part = part + (base * C.riv[e].share * h.ac);
    }
    double qb_tmp = warp_sum(part);
    bool use_max = true;          // state currently defined by calculate_qbmax
    double D_last = 0.0;
    if (qb_tmp <= mean_q) { qb_diff = 0.0; sol_flag = 2; }                    // :126-131
    int iter = 0;
    // every lane carries identical copies of the scalars (warp_sum returns the same bits to all lanes)
    while (fabs(qb_diff) > DCP_F32_1EM2) {                                    // :134
        if (++iter > DCP_INIT_MAX_ITER) { status |= DCP_ST_INIT_ITERCAP; break; }
        qb_tmp = calc_qb<false>(C, W, m, lane, D, meanatb, meanT0DT);
        use_max = false; D_last = D;
        qb_diff_prev = qb_diff;
        qb_diff = (qb_tmp - mean_q) * 1000.0;                                 // :140
        if (qb_diff > DCP_F32_1EM2) guess = guess - ndiff;
        else if (qb_diff < -DCP_F32_1EM2) guess = guess + ndiff;
        D = -meanSZM * log(guess / q0bar);                                    // :153
        if ((qb_diff_prev > 0.0) && (qb_diff < 0.0)) {                        // :156-172
            if (ndiff > DCP_F32_1EM9) {
                ndiff = ndiff / 10.0;
            } else {
                sol_flag = 0;
                if (fabs(qb_diff) < fabs(qb_diff_prev)) break;
                qb_tmp = calc_qb<false>(C, W, m, lane, D, meanatb, meanT0DT);
                D_last = D;
                break;
            }
        }
    }
    status |= (sol_flag & DCP_ST_SOLFLAG_MASK);

    // final state = that of the last calculate_qb / calculate_qbmax call
    if (use_max) {
        for (int ia = lane; ia < nac; ia += 32) {
            W.sd[ia * (size_t)B + m] = 0.0;
            W.qbf0[ia * (size_t)B + m] = W.q0[ia * (size_t)B + m];
        }
    } else {
        (void)calc_qb<true>(C, W, m, lane, D_last, meanatb, meanT0DT);
    }
    // srinit clip (:203-206): only layers some HRU references are touched
    for (int i = lane; i < npt; i += 32) {
        if (!C.layer_used[i]) continue;
        const double srmax = W.par[(2 * npt + i) * (size_t)B + m];
        if (W.par[(3 * npt + i) * (size_t)B + m] > srmax) {
            W.par[(3 * npt + i) * (size_t)B + m] = srmax;
            mp[3 * npt + i] = srmax;
        }
    }
    __syncwarp();
    for (int ia = lane; ia < nac; ia += 32) {                                 // :194-213
        const HruStatic h = C.hru[ia];
        const size_t o = ia * (size_t)B + m;
        const double qbf = W.qbf0[o];
        W.qint1[o] = qbf * h.ac;
        W.suz[o] = 0.0;
        const double srz = W.par[(2 * npt + h.ipar) * (size_t)B + m] - W.par[(3 * npt + h.ipar) * (size_t)B + m];
        W.srz[o] = srz;
        if (W.aet) W.aet[o] = 0.0;
        if (CHECK) {
            W.sum_pptin[o] = 0.0; W.sum_pein[o] = 0.0; W.sumexs[o] = 0.0; W.sumexus[o] = 0.0;
            W.sdini[o] = W.sd[o]; W.srzini[o] = srz; W.qbfini[o] = qbf;
        }
    }
    if (CHECK)
        for (int r = lane; r < 3 * C.nriv; r += 32) W.riv_sums[r * (size_t)B + m] = 0.0;
    for (int r = lane; r < C.nriv; r += 32) {
        W.sse[r * (size_t)B + m] = 0.0;
        W.lsse[r * (size_t)B + m] = 0.0;
        W.sq[r * (size_t)B + m] = 0.0; W.sqq[r * (size_t)B + m] = 0.0; W.sqo[r * (size_t)B + m] = 0.0;
        W.bad[r * (size_t)B + m] = 0;
    }
    if (lane == 0) {
        W.diag[0 * (size_t)B + m] = mean_q;
        W.diag[1 * (size_t)B + m] = qb_tmp;
        W.diag[2 * (size_t)B + m] = guess;
        W.status[m] = status;
    }
}

// ------------------------------------------------------------------------------------------------
// K2 (streamed): Topmod time loop (RRMODEL/dyna_topmod.f90:92-233) for steps [t0, t1)
// ------------------------------------------------------------------------------------------------
// Thread = member.  Per step the thread walks the HRUs in ascending order, exactly like the
// reference, so every accumulation (qin gather, qb, qof, the results_ts sums) has the reference's
// summation order and needs no cross-thread reduction.  qbf is double buffered in time
// (dynamic_dist reads last step's values of ALL HRUs before any is updated).
template <bool FAST, bool CHECK, bool AET, bool NPT1>
__global__ void __launch_bounds__(DCP_STREAM_THREADS)
k_physics_streamed(DevCatchment C, DevBatch W, int t0, int t1) {
    extern __shared__ double s_acc[];              // [2*nriv (+3*nriv)][blockDim] river accumulators
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    const int B = W.B;
    if (m >= B) return;
    const int nac = C.nac, nriv = C.nriv, npt = C.npt;
    const int nthr = blockDim.x;
    double *s_qb = s_acc + threadIdx.x;            // s_qb[r*nthr], s_qof[(nriv+r)*nthr]
    double *s_qof = s_acc + (size_t)nriv * nthr + threadIdx.x;

    StepConst g;
    g.dt = C.dt; g.dtt = C.dtt; g.inv_dtt = C.inv_dtt; g.ntt = C.ntt;

    // parameters of layer 0 hoisted when there is a single layer
    double p_szm = 0, p_srmax = 0, p_td = 0, p_smax = 0;
    if (NPT1) {
        p_szm = W.par[(0 * npt) * (size_t)B + m];
        p_srmax = W.par[(2 * npt) * (size_t)B + m];
        p_td = W.par[(5 * npt) * (size_t)B + m];
        p_smax = W.par[(6 * npt) * (size_t)B + m];
    }
    const double inv_szm = 1.0 / p_szm, inv_srmax = 1.0 / p_srmax;
    int status = 0;

    for (int t = t0; t < t1; ++t) {
        const double *__restrict__ qbf_old = (t & 1) ? W.qbf1 : W.qbf0;
        double *__restrict__ qbf_new = (t & 1) ? W.qbf0 : W.qbf1;
        for (int r = 0; r < nriv; ++r) { s_qb[r * nthr] = 0.0; s_qof[r * nthr] = 0.0; }
        const double *rain_t = C.rain + (size_t)t * C.ngp;
        const double *pet_t = C.pet + (size_t)t * C.nge;

        // results_ts running sums (dyna_results_ts.f90:56-71), CHECK only
        double c_qbf = 0, c_qbfini = 0, c_sd = 0, c_srz = 0, c_suz = 0, c_sdini = 0, c_srzini = 0,
               c_catin = 0, c_eout = 0, c_exs = 0, c_exus = 0;

        for (int ia = 0; ia < nac; ++ia) {
            const HruStatic h = C.hru[ia];
            const size_t o = ia * (size_t)B + m;
            const double qbf_o = qbf_old[o];
            // dynamic_dist, river part (dyna_dynamic_dist.f90:57-66): (((dtt*qbf)*w_riv)*rivers)*ac
            {
                const double base = C.dtt * qbf_o * h.w_riv;
                for (int e = h.riv_beg; e < h.riv_end; ++e) {
                    const RivEntry re = C.riv[e];
                    s_qb[re.r * nthr] = s_qb[re.r * nthr] + base * re.share * h.ac;
                }
            }
            // dynamic_dist, HRU part in gather form: sources ascending == reference's scatter order
            double qin = 0.0;
            for (int e = h.row_beg; e < h.row_end; ++e) {
                const WEntry we = C.w[e];
                qin = qin + qbf_old[we.src * (size_t)B + m] * we.w * we.ac_src;
            }
            HruConst c;
            double szm;
            if (NPT1) {
                szm = p_szm; c.srmax = p_srmax; c.td = p_td; c.smax = p_smax;
                c.inv_srmax = inv_srmax;
                c.inv_m2 = h.cosm * inv_szm;
            } else {
                szm = W.par[(0 * npt + h.ipar) * (size_t)B + m];
                c.srmax = W.par[(2 * npt + h.ipar) * (size_t)B + m];
                c.td = W.par[(5 * npt + h.ipar) * (size_t)B + m];
                c.smax = W.par[(6 * npt + h.ipar) * (size_t)B + m];
                if (FAST) { c.inv_srmax = 1.0 / c.srmax; c.inv_m2 = h.cosm / szm; }
            }
            c.ac = h.ac;
            if (FAST) c.inv_ac = 1.0 / h.ac;
            c.q0 = W.q0[o];
            c.q1 = c.q0 * h.cosm;
            c.q2 = W.q2[o];
            c.m2 = szm / h.cosm;
            c.ims_rz = h.ims_rz; c.ims_uz = h.ims_uz;

            HruState s;
            s.sd = W.sd[o]; s.suz = W.suz[o]; s.srz = W.srz[o]; s.qint1 = W.qint1[o];
            const double p = LDG(rain_t + h.ippt), ep = LDG(pet_t + h.ipet);
            HruFlux f;
            hru_step<FAST, CHECK>(c, g, p, ep, qin, s, f);

            W.sd[o] = s.sd; W.suz[o] = s.suz; W.srz[o] = s.srz; W.qint1[o] = s.qint1;
            qbf_new[o] = f.qbf;
            if (AET) { if (h.ims_rz == 1 && ep > 0.0) W.aet[o] = W.aet[o] + f.ea; }
            if (f.ex > 0.0) s_qof[h.ipriv * nthr] = s_qof[h.ipriv * nthr] + f.ex * h.ac;   // dyna_results_ac.f90:35-38

            if (CHECK) {
                if (f.rz_bad) status |= DCP_ST_RZ_BALANCE;
                double sp = W.sum_pptin[o], se = W.sum_pein[o], sx = W.sumexs[o], su = W.sumexus[o];
                if (h.ims_rz == 1) { sp = sp + p; if (ep > 0.0) se = se + ep; }
                if (f.exus > 0.0 || f.exus < 0.0) su = su + (f.exus * h.ac);
                if (f.exs1 != 0.0) sx = sx + (f.exs1 * h.ac);
                if (f.exs2 != 0.0) sx = sx + (f.exs2 * h.ac);
                W.sum_pptin[o] = sp; W.sum_pein[o] = se; W.sumexs[o] = sx; W.sumexus[o] = su;
                c_qbf = c_qbf + f.qbf * h.ac;
                c_qbfini = c_qbfini + (W.qbfini[o] * h.ac);
                c_sd = c_sd - (s.sd * h.ac);
                c_srz = c_srz + s.srz * h.ac;
                c_suz = c_suz + s.suz * h.ac;
                c_srzini = c_srzini + (W.srzini[o] * h.ac);
                c_sdini = c_sdini - (W.sdini[o] * h.ac);
                c_catin = c_catin + sp * h.ac;
                c_eout = c_eout + (AET ? W.aet[o] : 0.0) * h.ac;
                c_exs = c_exs + sx;
                c_exus = c_exus + su;
            }
        }
        // river (dyna_river.f90:42-49): qout = qb + qof
        double *qo = W.qout + (size_t)m * nriv * W.R + (t & W.Rmask);      // [member][river][time]
        double k_qb = 0, k_qof = 0, k_out = 0;
        for (int r = 0; r < nriv; ++r) {
            const double qb = s_qb[r * nthr], qof = s_qof[r * nthr];
            const double qout = qb + qof;
            qo[(size_t)r * W.R] = qout;
            if (CHECK) {
                double *rs = W.riv_sums + m;
                const double a = rs[(0 * nriv + r) * (size_t)B] + qb;     // sumqb (dyna_dynamic_dist.f90:63-65)
                const double b = rs[(1 * nriv + r) * (size_t)B] + qof;    // sumqof
                const double c2 = rs[(2 * nriv + r) * (size_t)B] + qout;  // sumqout
                rs[(0 * nriv + r) * (size_t)B] = a;
                rs[(1 * nriv + r) * (size_t)B] = b;
                rs[(2 * nriv + r) * (size_t)B] = c2;
                k_qb = k_qb + a; k_qof = k_qof + b; k_out = k_out + c2;
            }
        }
        if (CHECK) {   // dyna_results_ts.f90:73-101 (suzini = 0)
            const double sumallini = c_srzini + 0.0 + c_sdini;
            const double wbal = sumallini + c_catin - c_eout - k_out - (c_srz + c_suz + c_sd)
                                + ((c_qbfini - c_qbf) * C.dtt);
            if (fabs((c_exs + c_exus + k_qb) - k_out) > DCP_F32_1EM6) status |= DCP_ST_TS_OUTSUM;
            if (fabs(wbal) > DCP_F32_1EM6) status |= DCP_ST_TS_WBAL;
            (void)k_qof;
        }
    }
    if (CHECK && status) W.status[m] |= status;
}

// ------------------------------------------------------------------------------------------------
// Routing.  The reference keeps a flow matrix per node that is incremented by histogram*inflow and
// shifted one column per step (DTA/dta_route_processing.f90:671-681, 723-727).  Column
// node_hist_indexes(n,1) of row n therefore carries the FIR
//     y_n(t) = sum_k delay_hist_n(k) * qout_n(t-k),   accumulated oldest term first,
// and a column further downstream is the same series delayed.  Both facts make the routing
// time-parallel: one thread per (member, node, timestep), lanes along time so that the inflow window
// and the y series are read as coalesced, sliding lines ([member][node][time] layouts) and the
// histogram taps are broadcast.  The accumulation order per output equals the reference's.
// ------------------------------------------------------------------------------------------------
#define ROUTE_CH 256        // timesteps per block

// K3a: y_n(t) for t in [t0, t1)
__global__ void __launch_bounds__(ROUTE_CH)
k_route_conv(DevCatchment C, DevBatch W, int t0, int t1, int use_smem) {
    extern __shared__ double s_route[];              // taps[Lcap] | window[Lcap + ROUTE_CH]
    const int m = blockIdx.x, n = blockIdx.y;
    const int tb = t0 + blockIdx.z * ROUTE_CH;        // first step of this block
    const int t = tb + threadIdx.x;
    const int B = W.B;
    const int riv = C.node_input[n];
    const int bins = W.hbins[n * (size_t)B + m];
    const size_t R = W.R;
    double *yo = W.y + ((size_t)m * C.nnode + n) * R;
    if (riv < 0 || bins <= 0) {
        if (t < t1) yo[t & W.Rmask] = 0.0;
        return;
    }
    const double *dh = W.dh + ((size_t)m * C.nnode + n) * W.Lcap;
    const double *qi = W.qout + ((size_t)m * C.nriv + riv) * R;
    double y = 0.0;
    if (use_smem) {
        double *taps = s_route, *win = s_route + W.Lcap;
        for (int k = threadIdx.x; k < bins; k += ROUTE_CH) taps[k] = dh[k];
        // window element j holds qout(tb - (bins-1) + j)
        const int wlen = bins - 1 + ROUTE_CH;
        for (int j = threadIdx.x; j < wlen; j += ROUTE_CH) {
            const int tt = tb - (bins - 1) + j;
            win[j] = (tt >= 0 && tt < t1) ? qi[tt & W.Rmask] : 0.0;
        }
        __syncthreads();
        if (t < t1) {
            int k = bins - 1;
            if (k > t) k = t;                        // terms before the first timestep do not exist
            const double *wp = win + (bins - 1) + threadIdx.x;
            for (; k >= 0; --k) y = y + taps[k] * wp[-k];
        }
    } else if (t < t1) {
        int k = bins - 1;
        if (k > t) k = t;
        for (; k >= 0; --k) y = y + dh[k] * qi[(t - k) & W.Rmask];
    }
    if (t < t1) yo[t & W.Rmask] = y;
}

// flow_matrix(u, col) at step t (col = the gauge's sum_index)
__device__ __forceinline__ double node_at_column(const DevCatchment &C, const DevBatch &W, int m, int u,
                                                 int col, int t) {
    const int B = W.B;
    const int hs = W.hstart[u * (size_t)B + m];
    const int bins = W.hbins[u * (size_t)B + m];
    if (bins <= 0) return 0.0;
    const int d = hs - col;
    if (d >= 0) {                                    // column upstream of u's histogram: pure delay
        if (t - d < 0) return 0.0;
        return W.y[((size_t)m * C.nnode + u) * W.R + ((t - d) & W.Rmask)];
    }
    // column inside (or past) the histogram of u: partial sum, oldest first
    const int riv = C.node_input[u];
    if (riv < 0) return 0.0;
    const int o = -d;
    double y = 0.0;
    int k = bins - 1 - o;
    if (k > t) k = t;
    const double *dh = W.dh + ((size_t)m * C.nnode + u) * W.Lcap;
    const double *qi = W.qout + ((size_t)m * C.nriv + riv) * W.R;
    for (; k >= 0; --k) y = y + dh[k + o] * qi[(t - k) & W.Rmask];
    return y;
}

// K3b: q(iriv, t) = (own node + upstream tree at the gauge's column) / frac_sub_area
//      (DTA/dta_route_processing.f90:685-714, RRMODEL/dyna_river.f90:57-59) and the per-step objective terms.
__global__ void __launch_bounds__(ROUTE_CH)
k_route_gauge(DevCatchment C, DevBatch W, int t0, int t1, double *__restrict__ q_out /*[nriv x nstep x B] or null*/) {
    const int m = blockIdx.x, i = blockIdx.y;        // river / output index; flow_output_indexes(i) = i
    const int t = t0 + blockIdx.z * ROUTE_CH + threadIdx.x;
    if (t >= t1) return;
    const int B = W.B;
    const int col = W.hstart[i * (size_t)B + m];     // sum_index
    double q_sum = 0.0;
    q_sum = q_sum + node_at_column(C, W, m, i, col, t);
    for (int e = C.up_beg[i]; e < C.up_beg[i + 1]; ++e) q_sum = q_sum + node_at_column(C, W, m, C.up_idx[e], col, t);
    const double q = q_sum / C.frac_sub_area[i];
    if (q_out) q_out[((size_t)m * C.nstep + t) * C.nriv + i] = q;
    if (!isfinite(q)) W.bad[i * (size_t)B + m] = 1;
    if (W.fdc && t >= C.t_warm) atomicAdd(W.fdc + ((size_t)m * C.nriv + i) * DCP_FDC_BINS + dcp_fdc_bin(q), 1u);
    if (C.qobs != nullptr) {
        double d2 = 0.0, dl2 = 0.0;
        const int n_t = C.nstep < C.nstep_obs ? C.nstep : C.nstep_obs;
        if (t >= C.t_warm && t < n_t) {
            const double o = C.qobs[(size_t)t * C.nriv + i];
            if (o != C.flow_err) {
                const double eps = C.obs[i].eps;
                d2 = (q - o) * (q - o);
                const double dl = log(q > eps ? q : eps) - log(o > eps ? o : eps);
                dl2 = dl * dl;
            }
        }
        const size_t o2 = ((size_t)m * C.nriv + i) * W.R + (t & W.Rmask);
        W.d2[o2] = d2;
        W.dl2[o2] = dl2;
        W.qs[o2] = q;
    }
}

// K3c: sums of the objective terms in ascending time order.  Adding the zero of a masked step leaves a sum
// unchanged, so every result equals the oracle's sum over the valid steps only, term by term in the same order.
// One warp = 32 consecutive members of one gauge.  The three series of a member are time-fastest, so a tile of
// 32 members x 32 steps is read with coalesced 256-byte row loads into shared memory and then walked by
// lane = member along time (conflict-free: row pitch 33), five independent ordered chains per lane.
#define OBJ_T 32
__global__ void __launch_bounds__(32)
k_objective_sum(DevCatchment C, DevBatch W, int t0, int t1) {
    __shared__ double s_d2[32][OBJ_T + 1], s_dl2[32][OBJ_T + 1], s_q[32][OBJ_T + 1], s_o[OBJ_T];
    const int lane = threadIdx.x, i = blockIdx.y;
    const int m_base = blockIdx.x * 32, m = m_base + lane;
    const int rows = min(32, W.B - m_base);
    const int n_t = C.nstep < C.nstep_obs ? C.nstep : C.nstep_obs;
    const size_t a = i * (size_t)W.B + (m < W.B ? m : m_base);
    double sse = W.sse[a], lsse = W.lsse[a], sq = W.sq[a], sqq = W.sqq[a], sqo = W.sqo[a];
    for (int tb = t0; tb < t1; tb += OBJ_T) {
        const int t = tb + lane;
        const bool in = t < t1;
        const size_t col = (size_t)(t & W.Rmask);
        for (int r0 = 0; r0 < rows; r0 += 8) {                   // coalesced: lanes along time; 24 loads in flight
            double a[8], b[8], c[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool ok = in && r0 + u < rows;
                const size_t o = ((size_t)(m_base + min(r0 + u, rows - 1)) * C.nriv + i) * W.R + col;
                a[u] = ok ? W.d2[o] : 0.0;
                b[u] = ok ? W.dl2[o] : 0.0;
                c[u] = ok ? W.qs[o] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) { s_d2[r0 + u][lane] = a[u]; s_dl2[r0 + u][lane] = b[u]; s_q[r0 + u][lane] = c[u]; }
        }
        double ov = C.flow_err;                                  // masked steps carry the missing-data code
        if (in && t >= C.t_warm && t < n_t) ov = C.qobs[(size_t)t * C.nriv + i];
        s_o[lane] = ov;
        __syncwarp();
        if (lane < rows) {
#pragma unroll 8
            for (int k = 0; k < OBJ_T; ++k) {                    // lanes along members, ascending time
                sse = sse + s_d2[lane][k];
                lsse = lsse + s_dl2[lane][k];
                const double o = s_o[k];
                if (o != C.flow_err) {                           // moment sums over the valid steps only
                    const double q = s_q[lane][k];
                    sq = sq + q;
                    sqq = sqq + q * q;
                    sqo = sqo + q * o;
                }
            }
        }
        __syncwarp();
    }
    if (m < W.B) { W.sse[a] = sse; W.lsse[a] = lsse; W.sq[a] = sq; W.sqq[a] = sqq; W.sqo[a] = sqo; }
}

// ------------------------------------------------------------------------------------------------
// Standalone router (DTA/route_run_standalone.f90:118-189 -> route_processing_build_hist +
// route_process_run_timeseries, DTA/dta_route_processing.f90:583-630): one histogram set per series
// from its own velocity, the caller's inflow series laid into the routing layout, then K3a/K3b.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_route_prepare(DevCatchment C, DevBatch W, const double *__restrict__ v /*[B]*/) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= W.B) return;
    W.status[m] = build_hist_one(C, W, m, v[m]);
    for (int r = 0; r < C.nriv; ++r) W.bad[r * (size_t)W.B + m] = 0;
}

// inflow [nriv x nstep x B] (ABI layout, river fastest) -> qout [B][nriv][R] (time fastest)
__global__ void __launch_bounds__(ROUTE_CH)
k_route_load_inflow(DevCatchment C, DevBatch W, const double *__restrict__ inflow) {
    const int m = blockIdx.x, t = blockIdx.y * ROUTE_CH + threadIdx.x;
    if (t >= C.nstep) return;
    const double *src = inflow + ((size_t)m * C.nstep + t) * C.nriv;
    for (int r = 0; r < C.nriv; ++r) W.qout[((size_t)m * C.nriv + r) * W.R + t] = src[r];
}

__global__ void __launch_bounds__(128)
k_route_status(DevCatchment C, DevBatch W, int32_t *__restrict__ status) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= W.B) return;
    int st = W.status[m];
    for (int r = 0; r < C.nriv; ++r) if (W.bad[r * (size_t)W.B + m]) st |= DCP_ST_NONFINITE;
    status[m] = st;
}

// ------------------------------------------------------------------------------------------------
// K4: outputs in the ABI's layouts (member slowest)
// ------------------------------------------------------------------------------------------------
// Flow exceeded a fraction p of the time, from the histogram (counted from the top bin down, linear inside the bin)
__device__ double fdc_quantile(const uint32_t *__restrict__ h, double n_total, double p) {
    const double r = p * n_total;
    double cum = 0.0;
    for (int b = DCP_FDC_BINS - 1; b >= 0; --b) {
        const double cnt = (double)h[b];
        if (cnt == 0.0) continue;
        cum = cum + cnt;
        if (cum >= r) {
            if (b == 0) return 0.0;
            const int key = b - 1;                                                   // 8 * octave + sub-bin
            const double lo = ldexp(1.0 + (double)(key & 7) * 0.125, (key >> 3) + DCP_FDC_OCTAVE_LO);
            if (b == DCP_FDC_BINS - 1) return ldexp(1.0, DCP_FDC_OCTAVE_LO + DCP_FDC_OCTAVES);
            const double hi = ldexp(1.0 + (double)((key & 7) + 1) * 0.125, (key >> 3) + DCP_FDC_OCTAVE_LO);
            return lo + ((cum - r) / cnt) * (hi - lo);                              // (cum - r) / cnt of the bin lies below the target
        }
    }
    return 0.0;
}

__global__ void __launch_bounds__(128)
k_finalize(DevCatchment C, DevBatch W, double *__restrict__ perf, double *__restrict__ totals,
           double *__restrict__ diag, int32_t *__restrict__ status, double *__restrict__ sig) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    const int B = W.B;
    if (m >= B) return;
    const int nriv = C.nriv;
    int st = W.status[m];
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    for (int g = 0; g < nriv; ++g) {
        const int bad = W.bad[g * (size_t)B + m];
        if (bad) st |= DCP_ST_NONFINITE;
        if (perf) {
            double nse = nan, lnse = nan, kge = nan, pbias = nan, rmse = nan;
            if (C.qobs && !bad && C.obs[g].n > 0) {
                const ObsStat os = C.obs[g];
                const size_t a = g * (size_t)B + m;
                const double sse = W.sse[a], sq = W.sq[a];
                nse = 1.0 - sse / os.sst;
                lnse = 1.0 - W.lsse[a] / os.lsst;
                // KGE (Gupta et al. 2009) from the moment sums, PBIAS, RMSE: DESIGN.md "Objective"
                const double nn = (double)os.n;
                const double mean_s = sq / nn;
                const double var_s = W.sqq[a] / nn - mean_s * mean_s;
                const double var_o = os.sst / nn;
                const double cov = W.sqo[a] / nn - mean_s * os.mean;
                const double r = cov / sqrt(var_s * var_o);
                const double alpha = sqrt(var_s / var_o);
                const double beta = mean_s / os.mean;
                kge = 1.0 - sqrt(((r - 1.0) * (r - 1.0) + (alpha - 1.0) * (alpha - 1.0)) + (beta - 1.0) * (beta - 1.0));
                pbias = (100.0 * (sq - os.sum)) / os.sum;
                rmse = sqrt(sse / nn);
            }
            double *pf = perf + ((size_t)m * nriv + g) * DCP_NMEASURE;
            pf[DCP_PERF_NSE] = nse; pf[DCP_PERF_LOGNSE] = lnse; pf[DCP_PERF_KGE] = kge;
            pf[DCP_PERF_PBIAS] = pbias; pf[DCP_PERF_RMSE] = rmse;
        }
    }
    if (totals) {
        // dyna_write_output.f90:61-77: per reach sum(sum_x*ac)/sum(ac)*1000; ppt and pet are member independent
        for (int g = 0; g < nriv; ++g) {
            totals[((size_t)m * nriv + g) * 3 + 0] = (C.tot_ppt[g] / C.tot_frac[g]) * 1000.0;
            totals[((size_t)m * nriv + g) * 3 + 1] = (C.tot_pet[g] / C.tot_frac[g]) * 1000.0;
        }
        for (int g = 0; g < nriv; ++g) {
            double a = 0.0;
            for (int ia = 0; ia < C.nac; ++ia)
                if (C.hru[ia].ipriv == g) a = a + (W.aet[ia * (size_t)B + m] * C.hru[ia].ac);
            totals[((size_t)m * nriv + g) * 3 + 2] = (a / C.tot_frac[g]) * 1000.0;
        }
    }
    if (sig && W.fdc) {
        const double nan2 = __longlong_as_double(0x7ff8000000000000ll);
        for (int g = 0; g < nriv; ++g) {
            const uint32_t *h = W.fdc + ((size_t)m * nriv + g) * DCP_FDC_BINS;
            double n_total = 0.0;
            for (int b = 0; b < DCP_FDC_BINS; ++b) n_total = n_total + (double)h[b];
            double *sg = sig + ((size_t)m * nriv + g) * DCP_NSIG;
            const bool ok = n_total > 0.0 && !W.bad[g * (size_t)B + m];
            const double q33 = ok ? fdc_quantile(h, n_total, 0.33) : nan2, q66 = ok ? fdc_quantile(h, n_total, 0.66) : nan2;
            sg[DCP_SIG_Q5] = ok ? fdc_quantile(h, n_total, 0.05) : nan2;
            sg[DCP_SIG_Q50] = ok ? fdc_quantile(h, n_total, 0.50) : nan2;
            sg[DCP_SIG_Q95] = ok ? fdc_quantile(h, n_total, 0.95) : nan2;
            sg[DCP_SIG_SLOPE] = (ok && q66 > 0.0) ? (log(q33) - log(q66)) / 0.33 : nan2;
        }
    }
    if (diag)
        for (int k = 0; k < 3; ++k) diag[(size_t)m * 3 + k] = W.diag[k * (size_t)B + m];
    if (status) status[m] = st;
    W.status[m] = st;
}

// ------------------------------------------------------------------------------------------------
// FP64 pipe peak: 8 independent DFMA chains per thread, no memory traffic
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// device evaluation of the fast-math primitives (accuracy tests): op 0 exp, 1 log, 2 rcp, 3 table exp, 4 table log
__global__ void k_debug_fastmath(int op, int64_t n, const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ tab) {
    __shared__ double tab_s[FM_EXP_TAB + 2 * FM_LOG_TAB];
    for (int i = threadIdx.x; i < FM_EXP_TAB + 2 * FM_LOG_TAB; i += blockDim.x) tab_s[i] = tab[i];
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = x[i];
        y[i] = op == 0 ? fm_exp(v) : op == 1 ? fm_log(v) : op == 2 ? fm_rcp(v) : op == 3 ? fm_exp_tab(v, tab_s) : fm_log_tab(v, tab_s + FM_EXP_TAB);
    }
}

void dcp_launch_debug_fastmath(int op, int64_t n, const double *x, double *y, const double *tab, cudaStream_t s) {
    k_debug_fastmath<<<592, 256, 0, s>>>(op, n, x, y, tab);
}

// ------------------------------------------------------------------------------------------------
// launch wrappers (called from dcp_api.cu)
// ------------------------------------------------------------------------------------------------
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

void dcp_launch_min_chv(const double *mcpar, int npt, int64_t n_member, double *out, cudaStream_t s) {
    k_min_chv<<<148, 256, 0, s>>>(mcpar, npt, n_member, out);
}

void dcp_launch_init(const DevCatchment &C, const DevBatch &W, double *mcpar, bool check, cudaStream_t s) {
    const int grid = cdiv(W.B, 4);                       // one warp per member, 4 members per block
    if (check) k_init_members<true><<<grid, 128, 0, s>>>(C, W, mcpar);
    else k_init_members<false><<<grid, 128, 0, s>>>(C, W, mcpar);
}

// The river accumulators take 2*nriv doubles of shared memory per thread: the block shrinks (128 -> 64 -> 32 threads) until
// they fit the opt-in limit, so catchments with many gauges keep the balance checks; beyond ~450 gauges the launch is refused.
template <bool FAST, bool CHECK, bool AET>
static cudaError_t launch_streamed2(const DevCatchment &C, const DevBatch &W, int t0, int t1, cudaStream_t s) {
    int dev = 0, optin = 48 * 1024;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    int threads = DCP_STREAM_THREADS;
    while (threads > 32 && (size_t)2 * C.nriv * threads * sizeof(double) > (size_t)optin) threads >>= 1;
    const size_t smem = (size_t)2 * C.nriv * threads * sizeof(double);
    if (smem > (size_t)optin) return cudaErrorInvalidConfiguration;
    const int grid = cdiv(W.B, threads);
    cudaError_t e = cudaSuccess;
    if (C.npt == 1) {
        auto k = k_physics_streamed<FAST, CHECK, AET, true>;
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<grid, threads, smem, s>>>(C, W, t0, t1);
    } else {
        auto k = k_physics_streamed<FAST, CHECK, AET, false>;
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<grid, threads, smem, s>>>(C, W, t0, t1);
    }
    return cudaGetLastError();
}

cudaError_t dcp_launch_physics_streamed(const DevCatchment &C, const DevBatch &W, int t0, int t1, bool fast,
                                        bool check, bool aet, cudaStream_t s) {
    if (check) {            // checked mode always tracks aet (results_ts needs sum_aeout)
        return fast ? launch_streamed2<true, true, true>(C, W, t0, t1, s) : launch_streamed2<false, true, true>(C, W, t0, t1, s);
    } else if (aet) {
        return fast ? launch_streamed2<true, false, true>(C, W, t0, t1, s) : launch_streamed2<false, false, true>(C, W, t0, t1, s);
    }
    return fast ? launch_streamed2<true, false, false>(C, W, t0, t1, s) : launch_streamed2<false, false, false>(C, W, t0, t1, s);
}

cudaError_t dcp_launch_route(const DevCatchment &C, const DevBatch &W, int t0, int t1, double *q_out, cudaStream_t s) {
    const int nchunk = cdiv(t1 - t0, ROUTE_CH);
    const size_t smem = ((size_t)2 * W.Lcap + ROUTE_CH) * sizeof(double);
    const int use_smem = smem <= 96 * 1024;
    if (use_smem && smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(k_route_conv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dim3 g1(W.B, C.nnode, nchunk), g2(W.B, C.nriv, nchunk);      // members on grid.x (no 65535 limit)
    k_route_conv<<<g1, ROUTE_CH, use_smem ? smem : 0, s>>>(C, W, t0, t1, use_smem);
    k_route_gauge<<<g2, ROUTE_CH, 0, s>>>(C, W, t0, t1, q_out);
    if (C.qobs != nullptr) {
        dim3 g3(cdiv(W.B, 32), C.nriv);
        k_objective_sum<<<g3, 32, 0, s>>>(C, W, t0, t1);
    }
    return cudaGetLastError();
}

void dcp_launch_route_prepare(const DevCatchment &C, const DevBatch &W, const double *v, const double *inflow,
                              cudaStream_t s) {
    k_route_prepare<<<cdiv(W.B, 128), 128, 0, s>>>(C, W, v);
    dim3 g(W.B, cdiv(C.nstep, ROUTE_CH));
    k_route_load_inflow<<<g, ROUTE_CH, 0, s>>>(C, W, inflow);
}

void dcp_launch_route_status(const DevCatchment &C, const DevBatch &W, int32_t *status, cudaStream_t s) {
    k_route_status<<<cdiv(W.B, 128), 128, 0, s>>>(C, W, status);
}

void dcp_launch_finalize(const DevCatchment &C, const DevBatch &W, double *perf, double *totals, double *diag,
                         int32_t *status, double *sig, cudaStream_t s) {
    k_finalize<<<cdiv(W.B, 128), 128, 0, s>>>(C, W, perf, totals, diag, status, sig);
}

// ------------------------------------------------------------------------------------------------
// Behavioural filtering (dcp_select_behavioural): per-member pass mask + distance from the threshold, per-block counts,
// then an order-preserving scatter with host-scanned block offsets (<= 1024 blocks for 10^6 members).
// ------------------------------------------------------------------------------------------------
#define SEL_THREADS 1024
__device__ __forceinline__ bool sel_eval(const double *__restrict__ perf, int64_t m, int nriv, int measure, int gauge, double thr, int hib, double &dist) {
    bool pass = true;
    double worst = 1e300;
    for (int g = (gauge < 0 ? 0 : gauge); g < (gauge < 0 ? nriv : gauge + 1); ++g) {
        double v = perf[((size_t)m * nriv + g) * DCP_NMEASURE + measure];
        if (measure == DCP_PERF_PBIAS) v = fabs(v);
        const bool ok = hib ? (v >= thr) : (v <= thr);                   // NaN compares false
        pass = pass && ok;
        const double d = fabs(v - thr);
        worst = d < worst ? d : worst;
    }
    dist = worst;
    return pass;
}
__global__ void __launch_bounds__(SEL_THREADS)
k_select_count(const double *__restrict__ perf, int64_t n, int nriv, int measure, int gauge, double thr, int hib, int32_t *__restrict__ block_count) {
    const int64_t m = blockIdx.x * (int64_t)SEL_THREADS + threadIdx.x;
    double d;
    const bool pass = m < n && sel_eval(perf, m, nriv, measure, gauge, thr, hib, d);
    const int c = __syncthreads_count(pass);
    if (threadIdx.x == 0) block_count[blockIdx.x] = c;
}
__global__ void __launch_bounds__(SEL_THREADS)
k_select_scatter(const double *__restrict__ perf, int64_t n, int nriv, int measure, int gauge, double thr, int hib,
                 const int64_t *__restrict__ block_off, int64_t *__restrict__ index_out, double *__restrict__ dist_out) {
    __shared__ int warp_cnt[SEL_THREADS / 32];
    const int64_t m = blockIdx.x * (int64_t)SEL_THREADS + threadIdx.x;
    double d = 0.0;
    const bool pass = m < n && sel_eval(perf, m, nriv, measure, gauge, thr, hib, d);
    const unsigned bal = __ballot_sync(0xffffffffu, pass);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[w] = __popc(bal);
    __syncthreads();
    int base = 0;
    for (int i = 0; i < w; ++i) base += warp_cnt[i];
    if (pass) {
        const int64_t o = block_off[blockIdx.x] + base + __popc(bal & ((1u << lane) - 1u));
        index_out[o] = m;
        dist_out[o] = d;
    }
}
int dcp_select_blocks(int64_t n) { return (int)((n + SEL_THREADS - 1) / SEL_THREADS); }
void dcp_launch_select_count(const double *perf, int64_t n, int nriv, int measure, int gauge, double thr, int hib, int32_t *block_count, cudaStream_t s) {
    k_select_count<<<dcp_select_blocks(n), SEL_THREADS, 0, s>>>(perf, n, nriv, measure, gauge, thr, hib, block_count);
}
void dcp_launch_select_scatter(const double *perf, int64_t n, int nriv, int measure, int gauge, double thr, int hib, const int64_t *block_off,
                               int64_t *index_out, double *dist_out, cudaStream_t s) {
    k_select_scatter<<<dcp_select_blocks(n), SEL_THREADS, 0, s>>>(perf, n, nriv, measure, gauge, thr, hib, block_off, index_out, dist_out);
}

__global__ void k_or_status(int32_t *__restrict__ status, const int32_t *__restrict__ extra, int32_t mask, int64_t n) {
    const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (m < n) status[m] |= extra[m] & mask;
}
void dcp_launch_or_status(int32_t *status, const int32_t *extra, int32_t mask, int64_t n, cudaStream_t s) {
    k_or_status<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(status, extra, mask, n);
}

void dcp_launch_fp64_peak(double *out, int iters, int blocks, cudaStream_t s) {
    k_fp64_peak<<<blocks, 256, 0, s>>>(out, iters, 0.999999, 1e-9);
}
