// dcp_resident.cu — the persistent on-chip physics kernel (one launch per member batch).
//
// One CTA per SM, resident for the whole launch.  A CTA owns Mc ensemble members at a time and runs
// their ENTIRE simulation (all nstep timesteps) before it fetches the next member group:
//
//   * 512 compute threads: thread = one HRU (x G member sub-groups side by side when nac <= 256),
//     S member slots per thread.  The time-varying stores (sd, suz, srz, qint1[, aet]) live in
//     REGISTERS for the whole simulation; the per (HRU, member) constants (q0, q1, q2, m2, 1/m2) live
//     in shared memory in a [slot][thread] layout (each thread reads only its own column: no bank
//     conflicts, immediate offsets); the static HRU record is loaded into registers once.
//   * last step's baseflow qbf of every (member, HRU) lives in SHARED memory as S-tuples, double
//     buffered in time.  The dynamic_dist gather  qin(dest) = sum_src (qbf(src)*w)*ac(src)  walks a CSR
//     copy of the weighting matrix in shared memory, sources ascending (== the reference's scatter
//     order), each W entry loaded once and applied to the thread's S member slots: an SpMM with the
//     members as the dense dimension, fed by 32-byte vector loads of the qbf tuple.
//   * rain / PET tiles of Tt timesteps are staged into shared memory with TMA bulk copies
//     (cp.async.bulk + mbarrier, double buffered) by one elected thread and read by every member.
//   * "river" warps do the per-river sums off the critical path: lane = one summation chain
//     (member, river, qb | qof) walking the river's HRU list (kept in shared memory) in ascending HRU
//     order — bit-for-bit the reference's summation order — one step behind the compute warps; the
//     qb/qof lanes of a pair meet through one warp shuffle and stream qout(t) to global memory for the
//     routing kernels.  One __syncthreads per timestep is the only CTA-wide synchronisation.
//
// Hot path citations: RRMODEL/dyna_topmod.f90:92-233, dyna_dynamic_dist.f90:47-67, dyna_river.f90:42-49.
#include "dcp_device.cuh"
#include "dcp_physics.cuh"
#include "dcp_kernels.h"
#include "dcp_resident_common.cuh"
#include "dcp_route_fused.cuh"

#include <cstdint>

#define RES_COLS DCP_RES_COMPUTE      // unit columns per member slot (= HRU positions across the CTA)
#define RES_RIVER 128                 // one warpgroup of river warps
#define RES_ROUTE 32                  // H == 1 kernels: the LAST river warp routes finished tiles and adds up the objective terms
                                      // (dcp_route_fused.cuh) instead of summing; it is not part of the per-step CTA barrier.
                                      // (No extra warp: setmaxnreg trades registers inside the CTA's launch allocation only,
                                      // 640 x 96 = 512 x 112 + 128 x 32, so the thread count and the split must stay.)
#define RES_SUM(H) ((H) == 1 ? RES_RIVER - RES_ROUTE : RES_RIVER)   // lanes that run the summation chains
// hand-over river warps -> routing warp: once per DCP_RT_TILE steps and per member group
__device__ __forceinline__ void res_ab_sync() { asm volatile("bar.sync 2, %0;" ::"n"(RES_RIVER) : "memory"); }
#define RES_MAXPASS 4                 // chains per river lane
#define RES_PARW 6                    // srmax, td, smax, szm, 1/srmax, 1/szm

// ---------------------------------------------------------------------------------------------
// Shared-memory views and per-launch constants common to both warp roles
// ---------------------------------------------------------------------------------------------
struct ResSmem {
    double *qbf_s, *ex_s, *cst_s, *par_s, *w_s, *ac_s, *forc_s, *rqc_s;
    uint16_t *src_s, *rq_s, *ro_s;
    uint64_t *bars;
};

#define gc (StepConst{C.dt, C.dtt, C.inv_dtt, C.ntt})     /* straight from the constant bank, no registers */

// ---------------------------------------------------------------------------------------------
// compute role: 16 warps, thread = HRU x S member slots
// MODE 0: EXACT arithmetic (reference order, IEEE division, library exp/log)
// MODE 1: FAST arithmetic, generic branchy step (any ims flags / ntt)
// MODE 2: FAST arithmetic, LEAN branch-free step (dcp_physics.cuh: hru_step_lean)
// ---------------------------------------------------------------------------------------------
template <int S, int H, int LG, int MODE, bool AET>
__device__ __forceinline__ void res_compute_role(const DevCatchment &C, const DevBatch &W, const ResPlan &P,
                                                 const ResSmem &M) {
    constexpr bool FAST = MODE != 0;
    constexpr bool LEAN = MODE == 2;
    constexpr int CT = RES_COLS / H;                        // compute threads; each owns H HRU columns x S member slots
    constexpr int THREADS = CT + RES_SUM(H);               // threads at the per-step CTA barrier
    const int tid = threadIdx.x;
    const int nac_pad = P.nac_pad, Mc = P.Mc, B = W.B, npt = C.npt;
    const int qbf_stride = P.G * nac_pad * S;              // doubles per qbf time buffer
    const int forc_stride = P.Tt * (C.ngp + C.nge);

    // per-thread static HRU records (registers for the whole launch)
    int ia[H], gq[H], ippt[H], ipet[H], row_beg[H], row_end[H], my_tuple[H], g_tuple[H];
    double ac[H], inv_ac[H];
    const double *my_par[H];
#pragma unroll
    for (int j = 0; j < H; ++j) {
        const int col = tid + CT * j;                       // unit column 0..511
        const int pos = col % nac_pad;                      // position inside the member sub-group
        gq[j] = col / nac_pad;
        ia[j] = P.thread_hru[pos];                          // positions are sorted by W row length
        HruStatic h;
        if (ia[j] >= 0) h = C.hru[ia[j]];
        else { h = HruStatic{}; h.ac = 1.0; h.cosm = 1.0; }
        ac[j] = h.ac; inv_ac[j] = 1.0 / h.ac;
        ippt[j] = h.ippt; ipet[j] = h.ipet; row_beg[j] = h.row_beg; row_end[j] = h.row_end;
        my_tuple[j] = (gq[j] * nac_pad + pos) * S;          // offset of this column's qbf tuple
        g_tuple[j] = gq[j] * nac_pad * S;
        my_par[j] = M.par_s + ((size_t)gq[j] * S * npt + h.ipar) * RES_PARW;
    }
    const int par_kstride = npt * RES_PARW;
    const int ntiles = (C.nstep + P.Tt - 1) / P.Tt;
    uint32_t tiles_done = 0;                                // forcing tiles consumed so far (mbarrier phase tracking)
    long long busy = 0;

    for (int gi = blockIdx.x; gi < P.n_groups; gi += gridDim.x) {
        const int m_base = gi * Mc;
        cta_sync<THREADS>();                                // previous group fully done with shared memory

        // ---- load the group's initial state (written by k_init_members) ----
        double sd[H][S], suz[H][S], srz[H][S], qint1[H][S], aet[H][S];
        // the fat-thread variant (H == 2, 232 registers) also keeps the per-unit constants in registers
        constexpr bool CREG = false;   // tried: per-unit constants in registers for H == 2 -> 24 % slower (spills), kept in shared memory
        double rq0[H][S], rq1[H][S], rq2[H][S], rm2[H][S], rim2[H][S];
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const int col = tid + CT * j;
            const HruStatic h = ia[j] >= 0 ? C.hru[ia[j]] : HruStatic{};
#pragma unroll
            for (int k = 0; k < S; ++k) {
                const int m = m_base + gq[j] * S + k;
                sd[j][k] = 1.0; suz[j][k] = 0.0; srz[j][k] = 0.0; qint1[j][k] = 0.0; aet[j][k] = 0.0;
                double qb0 = 0.0, c_q0 = 1.0, c_q2 = 1.0, c_m2 = 1.0, c_q1 = 1.0, c_im2 = 1.0;
                if (ia[j] >= 0 && m < B) {
                    const size_t o = (size_t)ia[j] * B + m;
                    sd[j][k] = W.sd[o]; suz[j][k] = W.suz[o]; srz[j][k] = W.srz[o]; qint1[j][k] = W.qint1[o];
                    c_q0 = W.q0[o]; c_q2 = W.q2[o];
                    const double szm = W.par[(0 * npt + h.ipar) * (size_t)B + m];
                    c_m2 = szm / h.cosm;                    // dyna_satzone_analyticalsolver.f90:52
                    c_q1 = c_q0 * h.cosm;                   // :49
                    c_im2 = h.cosm / szm;
                    qb0 = W.qbf0[o];
                }
                M.qbf_s[my_tuple[j] + k] = qb0;             // time buffer 0 = step 0's "old" qbf
                M.qbf_s[qbf_stride + my_tuple[j] + k] = 0.0;
                M.cst_s[(0 * S + k) * RES_COLS + col] = c_q0;
                M.cst_s[(1 * S + k) * RES_COLS + col] = c_q2;
                M.cst_s[(2 * S + k) * RES_COLS + col] = c_m2;
                M.cst_s[(3 * S + k) * RES_COLS + col] = c_q1;
                M.cst_s[(4 * S + k) * RES_COLS + col] = c_im2;
                rq0[j][k] = c_q0; rq1[j][k] = c_q1; rq2[j][k] = c_q2; rm2[j][k] = c_m2; rim2[j][k] = c_im2;
                M.ex_s[(0 * S + k) * RES_COLS + col] = 0.0;
                M.ex_s[(1 * S + k) * RES_COLS + col] = 0.0;
            }
        }
        // member parameters -> shared (one thread per (member, layer))
        for (int i = tid; i < Mc * npt; i += CT) {
            const int jm = i / npt, l = i % npt;
            const int m = m_base + jm;
            double *pp = M.par_s + (size_t)i * RES_PARW;
            if (m < B) {
                const double srmax = W.par[(2 * npt + l) * (size_t)B + m];
                const double szm = W.par[(0 * npt + l) * (size_t)B + m];
                pp[0] = srmax;
                pp[1] = W.par[(5 * npt + l) * (size_t)B + m];
                pp[2] = W.par[(6 * npt + l) * (size_t)B + m];
                pp[3] = szm;
                pp[4] = 1.0 / srmax;
                pp[5] = 1.0 / szm;
            } else {
                pp[0] = pp[1] = pp[2] = pp[3] = pp[4] = pp[5] = 1.0;
            }
        }
        cta_sync<THREADS>();

        int t = 0;
        for (int tile = 0; tile < ntiles; ++tile, ++tiles_done) {
            const uint32_t b = tiles_done & 1;
            mbar_wait(&M.bars[b], (tiles_done >> 1) & 1);   // forcing tile landed (TMA, issued by the river role)
            const double *forc_t = M.forc_s + b * forc_stride;
            const int t_end = min(C.nstep, t + P.Tt);
            for (int tt = 0; t < t_end; ++t, ++tt) {
                const long long c0 = P.dbg ? clock64() : 0;
                const double *qbf_cur = M.qbf_s + (t & 1) * qbf_stride;
                double *qbf_nxt = M.qbf_s + ((t & 1) ^ 1) * qbf_stride;
                // ---- dynamic_dist gather (dyna_dynamic_dist.f90:49-52) for the thread's H rows ----
                double qin[H][S];
#pragma unroll
                for (int j = 0; j < H; ++j) {
#pragma unroll
                    for (int k = 0; k < S; ++k) qin[j][k] = 0.0;
                    const double *qg = qbf_cur + g_tuple[j];
                    for (int e = row_beg[j]; e < row_end[j]; ++e) {
                        const int sp = M.src_s[e];
                        const double w = M.w_s[e];
                        const double2 *qv = reinterpret_cast<const double2 *>(qg + sp * S);
                        double qs[S];
#pragma unroll
                        for (int k2 = 0; k2 < S / 2; ++k2) { const double2 v = qv[k2]; qs[2 * k2] = v.x; qs[2 * k2 + 1] = v.y; }
                        if (FAST) {
#pragma unroll
                            for (int k = 0; k < S; ++k) qin[j][k] = fma(qs[k], w, qin[j][k]);
                        } else {
                            const double acs = M.ac_s[sp];
#pragma unroll
                            for (int k = 0; k < S; ++k) qin[j][k] = qin[j][k] + qs[k] * w * acs;
                        }
                    }
                }
                // ---- root zone, unsaturated zone, saturated zone per (HRU, member slot) ----
                if (LEAN) {
                    // LG units per basic block: their dependency chains interleave (ILP)
#pragma unroll
                    for (int u0 = 0; u0 < H * S; u0 += LG) {
                        LeanOut lo[LG];
                        bool any_slow = false;
#pragma unroll
                        for (int v = 0; v < LG; ++v) {
                            const int j = (u0 + v) / S, k = (u0 + v) % S;
                            const int col = tid + CT * j;
                            const double *pp = my_par[j] + k * par_kstride;
                            const double *cst = M.cst_s + col;
                            HruConst c;
                            c.ac = ac[j]; c.inv_ac = inv_ac[j];
                            if (CREG) {
                                c.q0 = rq0[j][k]; c.q2 = rq2[j][k]; c.m2 = rm2[j][k]; c.q1 = rq1[j][k]; c.inv_m2 = rim2[j][k];
                            } else {
                                c.q0 = cst[(0 * S + k) * RES_COLS]; c.q2 = cst[(1 * S + k) * RES_COLS];
                                c.m2 = cst[(2 * S + k) * RES_COLS]; c.q1 = cst[(3 * S + k) * RES_COLS];
                                c.inv_m2 = cst[(4 * S + k) * RES_COLS];
                            }
                            c.srmax = pp[0]; c.td = pp[1]; c.smax = pp[2]; c.inv_srmax = pp[4];
                            c.ims_rz = 1; c.ims_uz = 1;
                            const double p = forc_t[tt * C.ngp + ippt[j]];
                            const double epp = fm_max(forc_t[P.Tt * C.ngp + tt * C.nge + ipet[j]], 0.0);
                            HruState st;
                            st.sd = sd[j][k]; st.suz = suz[j][k]; st.srz = srz[j][k]; st.qint1 = 0.0;
                            hru_step_lean(c, gc, p, epp, qin[j][k], st, lo[v]);
                            sd[j][k] = st.sd; suz[j][k] = st.suz; srz[j][k] = st.srz;     // qint1 is not carried: ntt == 1 => qin2 == qin/ac
                            any_slow |= lo[v].slow;
                        }
                        if (any_slow) {                                                    // rare: constants are re-read, not kept live
#pragma unroll
                            for (int v = 0; v < LG; ++v)
                                if (lo[v].slow) {
                                    const int j = (u0 + v) / S, k = (u0 + v) % S;
                                    const double *cst = M.cst_s + tid + CT * j;
                                    HruConst c;
                                    c.ac = ac[j]; c.inv_ac = inv_ac[j];
                                    c.q0 = cst[(0 * S + k) * RES_COLS]; c.q2 = cst[(1 * S + k) * RES_COLS];
                                    c.m2 = cst[(2 * S + k) * RES_COLS]; c.q1 = cst[(3 * S + k) * RES_COLS];
                                    c.inv_m2 = cst[(4 * S + k) * RES_COLS];
                                    c.smax = (my_par[j] + k * par_kstride)[2];
                                    hru_step_lean_fixup(c, gc, qin[j][k], qin[j][k], sd[j][k], lo[v]);
                                }
                        }
#pragma unroll
                        for (int v = 0; v < LG; ++v) {
                            const int j = (u0 + v) / S, k = (u0 + v) % S;
                            if (ia[j] >= 0) {
                                qbf_nxt[my_tuple[j] + k] = lo[v].qbf;
                                M.ex_s[((t & 1) * S + k) * RES_COLS + tid + CT * j] = lo[v].exac;
                            }
                            if (AET) aet[j][k] = aet[j][k] + lo[v].ea;
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < H; ++j) {
                        if (ia[j] < 0) continue;
                        const int col = tid + CT * j;
                        const HruStatic &hh = C.hru[ia[j]];
                        const int ims_rz = hh.ims_rz, ims_uz = hh.ims_uz;
                        const double p = forc_t[tt * C.ngp + ippt[j]];
                        const double ep = forc_t[P.Tt * C.ngp + tt * C.nge + ipet[j]];
                        const double *cst = M.cst_s + col;
#pragma unroll
                        for (int k = 0; k < S; ++k) {
                            const double *pp = my_par[j] + k * par_kstride;
                            HruConst c;
                            c.ac = ac[j]; c.inv_ac = inv_ac[j];
                            c.q0 = cst[(0 * S + k) * RES_COLS]; c.q2 = cst[(1 * S + k) * RES_COLS];
                            c.m2 = cst[(2 * S + k) * RES_COLS]; c.q1 = cst[(3 * S + k) * RES_COLS];
                            c.inv_m2 = cst[(4 * S + k) * RES_COLS];
                            c.srmax = pp[0]; c.td = pp[1]; c.smax = pp[2]; c.inv_srmax = pp[4];
                            c.ims_rz = ims_rz; c.ims_uz = ims_uz;
                            HruState s;
                            s.sd = sd[j][k]; s.suz = suz[j][k]; s.srz = srz[j][k]; s.qint1 = qint1[j][k];
                            HruFlux f;
                            hru_step<FAST, false>(c, gc, p, ep, qin[j][k], s, f);
                            sd[j][k] = s.sd; suz[j][k] = s.suz; srz[j][k] = s.srz; qint1[j][k] = s.qint1;
                            qbf_nxt[my_tuple[j] + k] = f.qbf;
                            M.ex_s[((t & 1) * S + k) * RES_COLS + col] = (f.ex > 0.0) ? f.ex * ac[j] : 0.0;   // dyna_results_ac.f90:35-38
                            if (AET) { if (ims_rz == 1 && ep > 0.0) aet[j][k] = aet[j][k] + f.ea; }
                        }
                    }
                }
                if (P.dbg) busy += clock64() - c0;
                cta_sync<THREADS>();
            }
        }
        if (P.dbg && (tid & 31) == 0) P.dbg[blockIdx.x * 20 + (tid >> 5)] = busy;
        if (AET) {
#pragma unroll
            for (int j = 0; j < H; ++j)
#pragma unroll
                for (int k = 0; k < S; ++k)
                    if (ia[j] >= 0 && m_base + gq[j] * S + k < B) W.aet[(size_t)ia[j] * B + (m_base + gq[j] * S + k)] = aet[j][k];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// river role: 4 warps.  Lane = summation chain; thread RES_CT also drives the forcing TMA.
// ---------------------------------------------------------------------------------------------
template <int S, int H, bool FAST>
__device__ __forceinline__ void res_river_role(const DevCatchment &C, const DevBatch &W, const ResPlan &P,
                                               const ResSmem &M) {
    constexpr int CT = RES_COLS / H;
    constexpr int THREADS = CT + RES_SUM(H);               // threads at the per-step CTA barrier
    constexpr bool FUSED = H == 1;                          // qout goes to the member slot's ring, a routing warp follows
    const int tid = threadIdx.x;
    const int nac_pad = P.nac_pad, Mc = P.Mc, B = W.B, nriv = C.nriv;
    const int qbf_stride = P.G * nac_pad * S;
    const int forc_stride = P.Tt * (C.ngp + C.nge);
    const int ntiles = (C.nstep + P.Tt - 1) / P.Tt;
    const uint32_t rain_bytes = (uint32_t)(P.Tt * C.ngp * sizeof(double));
    const uint32_t pet_bytes = (uint32_t)(P.Tt * C.nge * sizeof(double));
    uint32_t tiles_done = 0;
    const int rlane = tid - CT;
    constexpr int n_rlanes = RES_SUM(H);
    const int n_chains = 2 * Mc * nriv;                     // chain id = (pair << 1) | type, pair = member * nriv + river
    // EXACT: one lane per chain, terms added in ascending HRU order (the reference's order).
    // FAST : a team of T lanes per chain (strided partial sums + shuffle tree) -- order is free there.
    const int T = FAST ? P.river_team : 1;
    const int sub = rlane & (T - 1);
    const int chains_per_pass = n_rlanes / T;
    long long busy = 0;

    for (int gi = blockIdx.x; gi < P.n_groups; gi += gridDim.x) {
        const int m_base = gi * Mc;
        if (FUSED) res_ab_sync();                       // the routing warp is done with the rings of the previous group
        cta_sync<THREADS>();
        if (tid == CT) {                                // forcing tile 0 of this group
            const uint32_t b = tiles_done & 1;
            fence_proxy_async();
            mbar_expect_tx(&M.bars[b], rain_bytes + pet_bytes);
            tma_load_1d(M.forc_s + b * forc_stride, C.rain, rain_bytes, &M.bars[b]);
            tma_load_1d(M.forc_s + b * forc_stride + P.Tt * C.ngp, C.pet, pet_bytes, &M.bars[b]);
        }
        cta_sync<THREADS>();

        double qb_prev[RES_MAXPASS];
#pragma unroll
        for (int q = 0; q < RES_MAXPASS; ++q) qb_prev[q] = 0.0;

        int t = 0;
        for (int tile = 0; tile < ntiles; ++tile, ++tiles_done) {
            if (tid == CT && tile + 1 < ntiles) {
                // prefetch the next tile into the other buffer (everyone left it at the last barrier)
                const uint32_t nb = (tiles_done & 1) ^ 1;
                const size_t t_next = (size_t)(tile + 1) * P.Tt;
                fence_proxy_async();
                mbar_expect_tx(&M.bars[nb], rain_bytes + pet_bytes);
                tma_load_1d(M.forc_s + nb * forc_stride, C.rain + t_next * C.ngp, rain_bytes, &M.bars[nb]);
                tma_load_1d(M.forc_s + nb * forc_stride + P.Tt * C.ngp, C.pet + t_next * C.nge, pet_bytes, &M.bars[nb]);
            }
            const int t_end = min(C.nstep, t + P.Tt);
            for (; t <= t_end; ++t) {
                // step t: qb(t) chains from qbf_cur, qof(t-1) chains from last step's ex terms.
                // t == nstep (only after the last tile) drains qof(nstep-1).
                if (t == t_end && t < C.nstep) break;
                const bool drain = (t == C.nstep);
                const long long c0 = P.dbg ? clock64() : 0;
                const double *qbf_cur = M.qbf_s + (t & 1) * qbf_stride;
                const double *ex_prev = M.ex_s + ((t & 1) ^ 1) * (S * RES_COLS);
#pragma unroll
                for (int q = 0; q < RES_MAXPASS; ++q) {
                    if (q * chains_per_pass < n_chains) {                 // uniform over the river warps
                        const int cid = rlane / T + q * chains_per_pass;
                        const bool live = cid < n_chains;
                        const int type = cid & 1, pair = cid >> 1;
                        const int jm = live ? pair / nriv : 0, r = live ? pair - jm * nriv : 0;
                        const int gg = jm / S, kk = jm - gg * S;
                        // both chain kinds run the same loop:  acc += (((f0*v)*c0)*c1)*c2  with unit factors for qof
                        const uint16_t *lst; const double *val; int vstride, cstride, beg, end; double f0;
                        if (type == 0) {
                            lst = M.rq_s; beg = P.rivq_beg[r]; end = P.rivq_beg[r + 1];
                            val = qbf_cur + gg * nac_pad * S + kk; vstride = S; cstride = 1; f0 = FAST ? 1.0 : C.dtt;
                            if (drain) end = beg;
                        } else {
                            lst = M.ro_s; beg = P.rivo_beg[r]; end = P.rivo_beg[r + 1];
                            val = ex_prev + kk * RES_COLS + gg * nac_pad; vstride = 1; cstride = 0; f0 = 1.0;
                        }
                        if (!live) end = beg;
                        const double *cf = M.rqc_s + (cstride ? 0 : P.nq);
                        double acc = 0.0;
                        int i = beg + sub;
                        for (; i + 3 * T < end; i += 4 * T) {
                            double v[4], a0[4], a1[4], a2[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int iu = i + u * T;
                                v[u] = val[lst[iu] * vstride];
                                a0[u] = cf[iu * cstride];
                                if (!FAST) { a1[u] = cf[(P.nq + 1) + iu * cstride]; a2[u] = cf[2 * (P.nq + 1) + iu * cstride]; }
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                acc = FAST ? fma(v[u], a0[u], acc) : acc + f0 * v[u] * a0[u] * a1[u] * a2[u];   // dyna_dynamic_dist.f90:59-61
                        }
                        for (; i < end; i += T) {
                            const double vv = val[lst[i] * vstride];
                            const double b0 = cf[i * cstride];
                            if (FAST) acc = fma(vv, b0, acc);
                            else acc = acc + f0 * vv * b0 * cf[(P.nq + 1) + i * cstride] * cf[2 * (P.nq + 1) + i * cstride];
                        }
                        if (FAST)
                            for (int o = T >> 1; o > 0; o >>= 1) acc = acc + __shfl_xor_sync(0xffffffffu, acc, o);
                        const double qof = __shfl_down_sync(0xffffffffu, acc, T);           // the pair's qof(t-1) team
                        if (type == 0 && live && sub == 0) {
                            const int m = m_base + jm;
                            if (t > 0 && m < B)
                                W.qout[((size_t)(FUSED ? blockIdx.x * Mc + jm : m) * nriv + r) * W.R + ((t - 1) & W.Rmask)] = qb_prev[q] + qof;   // dyna_river.f90:43
                        }
                        if (type == 0) qb_prev[q] = acc;
                    }
                }
                if (P.dbg) busy += clock64() - c0;
                if (FUSED && t > 0 && (((t - 1) & (DCP_RT_TILE - 1)) == DCP_RT_TILE - 1 || drain)) res_ab_sync();   // tile of qout complete
                if (drain) { ++t; break; }
                cta_sync<THREADS>();
            }
        }
        if (P.dbg && (tid & 31) == 0) P.dbg[blockIdx.x * 20 + (tid >> 5)] = busy;
    }
}

// ---------------------------------------------------------------------------------------------
// H = HRU columns per compute thread.  H = 1: 16 compute warps x 112 registers.  H = 2: 8 compute warps
// x 232 registers, 8 (HRU, member) units per thread of which LG run interleaved in one basic block —
// fewer, fatter threads trade thread-level for instruction-level parallelism on the FP64 pipe.
template <int H> struct ResRegs;
template <> struct ResRegs<1> { static constexpr int compute = 112, river = 32; };   // launch cap 96: (96-32)*128 == (112-96)*512
template <> struct ResRegs<2> { static constexpr int compute = 232, river = 40; };   // launch cap 168: (168-40)*128 == (232-168)*256

// routing warp of the H == 1 kernels: one tile behind the river warps (dcp_route_fused.cuh)
template <bool FAST>
__device__ __forceinline__ void res_route_role(const DevCatchment &C, const DevBatch &W, const ResPlan &P) {
    for (int gi = blockIdx.x; gi < P.n_groups; gi += gridDim.x) {
        res_ab_sync();                                  // group start (pairs with the river warps')
        for (int t0 = 0; t0 < C.nstep; t0 += DCP_RT_TILE) {
            res_ab_sync();                              // qout of [t0, t0 + tile) is in the rings
            rt_tile<FAST>(C, W, gi * P.Mc, P.Mc, blockIdx.x * P.Mc, t0, min(C.nstep, t0 + DCP_RT_TILE));
        }
    }
}

template <int S, int H, int LG, int MODE, bool AET>
__global__ void __launch_bounds__(RES_COLS / H + RES_RIVER, 1)
k_physics_resident(DevCatchment C, DevBatch W, ResPlan P) {
    constexpr bool FAST = MODE != 0;
    constexpr int CT = RES_COLS / H;
    constexpr int NTHREADS = CT + RES_RIVER;                // threads of the CTA
    constexpr int THREADS = CT + RES_SUM(H);               // threads at the per-step CTA barrier
    extern __shared__ __align__(128) unsigned char smem[];
    ResSmem M;
    M.qbf_s = (double *)(smem + P.off_qbf);          // [2][G*nac_pad][S]   qbf tuples, time double buffer
    M.ex_s = (double *)(smem + P.off_ex);            // [2][S][512]         ex*ac terms
    M.cst_s = (double *)(smem + P.off_cst);          // [5][S][512]         q0, q2, m2, q1, 1/m2
    M.par_s = (double *)(smem + P.off_par);          // [Mc][npt][RES_PARW]
    M.w_s = (double *)(smem + P.off_w);              // [nnz]    w (EXACT) or w*ac_src (FAST)
    M.src_s = (uint16_t *)(smem + P.off_src);        // [nnz]    thread position of the source HRU
    M.ac_s = (double *)(smem + P.off_ac);            // [nac_pad] ac by thread position
    M.forc_s = (double *)(smem + P.off_forc);        // [2][Tt*(ngp+nge)]
    M.bars = (uint64_t *)(smem + P.off_bar);         // [2]
    M.rqc_s = (double *)(smem + P.off_rqc);          // [3][nq+1] river-term coefficients; [.][nq] = 1.0
    M.rq_s = (uint16_t *)(smem + P.off_rq);          // [nq]  thread positions, per river ascending HRU (share != 0)
    M.ro_s = (uint16_t *)(smem + P.off_ro);          // [nac] thread positions, per river ascending HRU (ipriv == r)

    const int tid = threadIdx.x;
    // ---- one-time CTA setup ----
    for (int e = tid; e < P.nnz; e += NTHREADS) {
        const WEntry we = C.w[e];
        M.w_s[e] = FAST ? we.w * we.ac_src : we.w;
        M.src_s[e] = (uint16_t)P.hru_pos[we.src];
    }
    for (int i = tid; i < P.nac_pad; i += NTHREADS) {
        const int hh = P.thread_hru[i];
        M.ac_s[i] = hh >= 0 ? C.hru[hh].ac : 0.0;
    }
    for (int i = tid; i < P.nq; i += NTHREADS) {
        const RivChain rc = P.rivq[i];
        M.rq_s[i] = (uint16_t)P.hru_pos[rc.ia];
        if (FAST) {
            M.rqc_s[i] = C.dtt * rc.w_riv * rc.share * rc.ac;
        } else {
            M.rqc_s[i] = rc.w_riv; M.rqc_s[(P.nq + 1) + i] = rc.share; M.rqc_s[2 * (P.nq + 1) + i] = rc.ac;
        }
    }
    if (tid < 3) M.rqc_s[tid * (P.nq + 1) + P.nq] = 1.0;
    for (int i = tid; i < C.nac; i += NTHREADS) M.ro_s[i] = (uint16_t)P.hru_pos[P.rivo[i]];
    if (tid == CT) {
        mbar_init(&M.bars[0], 1);
        mbar_init(&M.bars[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    // ---- warp-role split with register reallocation (setmaxnreg) ----
    if (tid < CT) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(ResRegs<H>::compute));
        res_compute_role<S, H, LG, MODE, AET>(C, W, P, M);
    } else if (tid < THREADS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ResRegs<H>::river));
        res_river_role<S, H, FAST>(C, W, P, M);
    } else if (H == 1) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ResRegs<H>::river));
        res_route_role<FAST>(C, W, P);
    }
}
#undef gc

// ---------------------------------------------------------------------------------------------
#ifndef DCP_RES_LEAN_TU
// ---------------------------------------------------------------------------------------------
// This TU (built with -fmad=false) holds the EXACT and generic-FAST kernels; the LEAN kernels are
// instantiated in dcp_resident_lean.cu, which includes this file with FMA contraction enabled — the
// lean step has no operation-order contract, and contraction removes ~10 % of its FP64 instructions.
cudaError_t dcp_launch_physics_resident_lean(const DevCatchment &C, const DevBatch &W, const ResPlan &P, bool aet,
                                             int grid, cudaStream_t s);

template <int S, int H>
static cudaError_t launch_res(const DevCatchment &C, const DevBatch &W, const ResPlan &P, bool fast, bool aet,
                              int grid, cudaStream_t s) {
    void (*k)(DevCatchment, DevBatch, ResPlan);
    if (fast) k = aet ? k_physics_resident<S, H, 1, 1, true> : k_physics_resident<S, H, 1, 1, false>;
    else k = aet ? k_physics_resident<S, H, 1, 0, true> : k_physics_resident<S, H, 1, 0, false>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, P.smem_bytes);
    if (e != cudaSuccess) return e;
    k<<<grid, RES_COLS / H + RES_RIVER, P.smem_bytes, s>>>(C, W, P);
    return cudaGetLastError();
}

cudaError_t dcp_launch_physics_resident(const DevCatchment &C, const DevBatch &W, const ResPlan &P, bool fast,
                                        bool aet, int grid, cudaStream_t s) {
    if (fast && P.lean_ok) return dcp_launch_physics_resident_lean(C, W, P, aet, grid, s);
    // the branchy exact / generic-fast steps run fastest with 16 compute warps
    return launch_res<DCP_RES_S, 1>(C, W, P, fast, aet, grid, s);
}
#else
// ---------------------------------------------------------------------------------------------
// LEAN kernels (this translation unit is built with FMA contraction on)
template <int S, int H, int LG>
static cudaError_t launch_lean(const DevCatchment &C, const DevBatch &W, const ResPlan &P, bool aet, int grid, cudaStream_t s) {
    void (*k)(DevCatchment, DevBatch, ResPlan) = aet ? k_physics_resident<S, H, LG, 2, true> : k_physics_resident<S, H, LG, 2, false>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, P.smem_bytes);
    if (e != cudaSuccess) return e;
    k<<<grid, RES_COLS / H + RES_RIVER, P.smem_bytes, s>>>(C, W, P);
    return cudaGetLastError();
}

cudaError_t dcp_launch_physics_resident_lean(const DevCatchment &C, const DevBatch &W, const ResPlan &P, bool aet,
                                             int grid, cudaStream_t s) {
    // the branch-free step profits from fewer, fatter threads (8 units per thread, DCP_RES_LG2 interleaved)
    const int variant = P.variant ? P.variant : 2;
    if (variant == 2) return launch_lean<DCP_RES_S, 2, DCP_RES_LG2>(C, W, P, aet, grid, s);
    return launch_lean<DCP_RES_S, 1, DCP_RES_LG>(C, W, P, aet, grid, s);
}
#endif
