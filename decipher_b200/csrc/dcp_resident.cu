// dcp_resident.cu — the persistent on-chip physics kernel (one launch per member batch).
//
// One CTA per SM, resident for the whole launch.  A CTA owns Mc ensemble members at a time and runs
// their ENTIRE simulation (all nstep timesteps) before it fetches the next member group:
//
//   * 512 compute threads: thread = one HRU (x G member sub-groups side by side when nac <= 256),
//     S member slots per thread.  The HRU store state (sd, suz, srz, qint1[, aet]) and the per
//     (HRU, member) constants (q0, q2, m2) live in REGISTERS for the whole simulation; the static
//     HRU record (area, slope cosine, forcing columns, W row bounds) is loaded into registers once.
//   * last step's baseflow qbf of every (member, HRU) lives in SHARED memory, double buffered in
//     time; the dynamic_dist gather  qin(dest) = sum_src (qbf(src)*w)*ac(src)  walks a CSR copy of the
//     weighting matrix held in shared memory, sources ascending (== the reference's scatter order),
//     each W entry loaded once and applied to the thread's S member slots (SpMM with members as the
//     dense dimension, 32-byte vector loads of the S-member qbf tuple).
//   * rain / PET tiles of Tt timesteps are staged into shared memory with TMA bulk copies
//     (cp.async.bulk + mbarrier, double buffered) by one elected thread and read by every member.
//   * one extra "river" warp does the per-river sums off the critical path: lane = (member, river)
//     chain, summing qb (from the qbf tuple in shared memory) and qof (from the ex*ac terms the
//     compute threads publish) over the HRUs in ascending order — bit-for-bit the reference's
//     summation order — one step behind the compute warps, and streams qout(t) to global memory for
//     the routing kernels.  One __syncthreads per timestep is the only CTA-wide synchronisation.
//
// Hot path citations: RRMODEL/dyna_topmod.f90:92-233, dyna_dynamic_dist.f90:47-67, dyna_river.f90:42-49.
#include "dcp_device.cuh"
#include "dcp_physics.cuh"
#include "dcp_kernels.h"

#include <cstdint>

#define RES_COMPUTE 512
#define RES_THREADS (RES_COMPUTE + 32)
#define RES_MAXPASS 4                 // river chains per lane of the river warp
#define RES_PARW 6                    // srmax, td, smax, szm, 1/srmax, 1/szm

// ---------------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy helpers (sm_90+ PTX; SASS: UBLKCP, SYNCS)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// MODE 0: EXACT arithmetic (reference order, IEEE division, library exp/log)
// MODE 1: FAST arithmetic, generic branchy step (any ims flags / ntt)
// MODE 2: FAST arithmetic, LEAN branch-free step, two member slots interleaved per basic block
template <int S, int MODE, bool AET>
__global__ void __launch_bounds__(RES_THREADS, 1)
k_physics_resident(DevCatchment C, DevBatch W, ResPlan P) {
    constexpr bool FAST = MODE != 0;
    constexpr bool LEAN = MODE == 2;
    extern __shared__ __align__(128) unsigned char smem[];
    double *qbf_s = (double *)(smem + P.off_qbf);          // [2][G*nac_pad*S]
    double *ex_s = (double *)(smem + P.off_ex);            // [2][Mc*nac_pad]
    double *cst_s = (double *)(smem + P.off_cst);          // [5][G*nac_pad*S]  q0, q2, m2, q1, 1/m2 per (HRU, member)
    double *par_s = (double *)(smem + P.off_par);          // [Mc][npt][RES_PARW]
    double *w_s = (double *)(smem + P.off_w);              // [nnz]   w (EXACT) or w*ac_src (FAST)
    uint16_t *src_s = (uint16_t *)(smem + P.off_src);      // [nnz]
    double *ac_s = (double *)(smem + P.off_ac);            // [nac_pad]
    double *forc_s = (double *)(smem + P.off_forc);        // [2][Tt*(ngp+nge)]
    uint64_t *bars = (uint64_t *)(smem + P.off_bar);       // [2]

    const int tid = threadIdx.x;
    const bool is_compute = tid < RES_COMPUTE;
    const int nac_pad = P.nac_pad, Mc = P.Mc, B = W.B, nriv = C.nriv, npt = C.npt;
    const int qbf_stride = P.G * nac_pad * S;              // doubles per time buffer
    const int ex_stride = Mc * nac_pad;
    const int forc_stride = P.Tt * (C.ngp + C.nge);
    const int Rm = W.Rmask;

    // ---- one-time CTA setup: W copy to shared memory, barriers ----
    for (int e = tid; e < P.nnz; e += RES_THREADS) {
        const WEntry we = C.w[e];
        w_s[e] = FAST ? we.w * we.ac_src : we.w;
        src_s[e] = (uint16_t)we.src;
    }
    for (int i = tid; i < nac_pad; i += RES_THREADS) ac_s[i] = i < C.nac ? C.hru[i].ac : 0.0;
    if (tid == RES_COMPUTE) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }

    // ---- per-thread static HRU record (registers for the whole launch) ----
    const int g = is_compute ? tid / nac_pad : 0;
    const int ia = is_compute ? P.thread_hru[tid % nac_pad] : -1;      // permuted so warps see similar W row lengths
    HruStatic h;
    if (ia >= 0) h = C.hru[ia];
    else { h = HruStatic{}; h.ac = 1.0; h.cosm = 1.0; }
    const double inv_ac = 1.0 / h.ac;
    const int gbase = g * nac_pad;

#define gc (StepConst{C.dt, C.dtt, C.inv_dtt, C.ntt})     /* straight from the constant bank, no registers */
    const int ntiles = (C.nstep + P.Tt - 1) / P.Tt;
    const uint32_t rain_bytes = (uint32_t)(P.Tt * C.ngp * sizeof(double));
    const uint32_t pet_bytes = (uint32_t)(P.Tt * C.nge * sizeof(double));
    uint32_t tiles_done = 0;                               // forcing tiles consumed so far (barrier phase tracking)

    for (int gi = blockIdx.x; gi < P.n_groups; gi += gridDim.x) {
        const int m_base = gi * Mc;
        __syncthreads();                                   // previous group fully done with shared memory

        // ---- load the group's initial state (written by k_init_members) ----
        // time-varying stores stay in registers; the per (HRU, member) constants q0, q2, m2 go to shared
        // memory as S-tuples.  Slots beyond the batch (m >= B) run on benign dummy values: no predicates.
        double sd[S], suz[S], srz[S], qint1[S], aet[S];
        if (is_compute) {
#pragma unroll
            for (int k = 0; k < S; ++k) {
                const int m = m_base + g * S + k;
                sd[k] = 1.0; suz[k] = 0.0; srz[k] = 0.0; qint1[k] = 0.0; aet[k] = 0.0;
                double qb0 = 0.0, c_q0 = 1.0, c_q2 = 1.0, c_m2 = 1.0, c_q1 = 1.0, c_im2 = 1.0;
                if (ia >= 0 && m < B) {
                    const size_t o = (size_t)ia * B + m;
                    sd[k] = W.sd[o]; suz[k] = W.suz[o]; srz[k] = W.srz[o]; qint1[k] = W.qint1[o];
                    c_q0 = W.q0[o]; c_q2 = W.q2[o];
                    const double szm = W.par[(0 * npt + h.ipar) * (size_t)B + m];
                    c_m2 = szm / h.cosm;                   // dyna_satzone_analyticalsolver.f90:52
                    c_q1 = c_q0 * h.cosm;                  // :49
                    c_im2 = h.cosm / szm;
                    qb0 = W.qbf0[o];
                }
                if (ia >= 0) {
                    qbf_s[(gbase + ia) * S + k] = qb0;     // time buffer 0 = step 0's "old" qbf
                    cst_s[0 * qbf_stride + (gbase + ia) * S + k] = c_q0;
                    cst_s[1 * qbf_stride + (gbase + ia) * S + k] = c_q2;
                    cst_s[2 * qbf_stride + (gbase + ia) * S + k] = c_m2;
                    cst_s[3 * qbf_stride + (gbase + ia) * S + k] = c_q1;
                    cst_s[4 * qbf_stride + (gbase + ia) * S + k] = c_im2;
                }
            }
            if (ia < 0) {                                   // idle thread: zero its padding rows in both buffers
                const int ip = tid % nac_pad;
                if (ip >= C.nac)
                    for (int k = 0; k < S; ++k) {
                        qbf_s[(gbase + ip) * S + k] = 0.0;
                        qbf_s[qbf_stride + (gbase + ip) * S + k] = 0.0;
                    }
            }
            // member parameters -> shared (one thread per (member, layer))
            for (int i = tid; i < Mc * npt; i += RES_COMPUTE) {
                const int jm = i / npt, l = i % npt;
                const int m = m_base + jm;
                double *pp = par_s + (size_t)i * RES_PARW;
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
        } else if (tid == RES_COMPUTE) {
            // forcing tile 0 of this group
            const uint32_t b = tiles_done & 1;
            fence_proxy_async();
            mbar_expect_tx(&bars[b], rain_bytes + pet_bytes);
            tma_load_1d(forc_s + b * forc_stride, C.rain, rain_bytes, &bars[b]);
            tma_load_1d(forc_s + b * forc_stride + P.Tt * C.ngp, C.pet, pet_bytes, &bars[b]);
        }
        __syncthreads();

        // river warp: chains (member jm, river r)
        double qb_prev[RES_MAXPASS];
#pragma unroll
        for (int q = 0; q < RES_MAXPASS; ++q) qb_prev[q] = 0.0;

        int t = 0;
        for (int tile = 0; tile < ntiles; ++tile, ++tiles_done) {
            const uint32_t b = tiles_done & 1;
            if (tid == RES_COMPUTE && tile + 1 < ntiles) {
                // prefetch the next tile into the other buffer (everyone left it at the last __syncthreads)
                const uint32_t nb = b ^ 1;
                const size_t t_next = (size_t)(tile + 1) * P.Tt;
                fence_proxy_async();
                mbar_expect_tx(&bars[nb], rain_bytes + pet_bytes);
                tma_load_1d(forc_s + nb * forc_stride, C.rain + t_next * C.ngp, rain_bytes, &bars[nb]);
                tma_load_1d(forc_s + nb * forc_stride + P.Tt * C.ngp, C.pet + t_next * C.nge, pet_bytes, &bars[nb]);
            }
            if (is_compute) mbar_wait(&bars[b], (tiles_done >> 1) & 1);
            const double *rain_t = forc_s + b * forc_stride;
            const double *pet_t = rain_t + P.Tt * C.ngp;
            const int t_end = min(C.nstep, t + P.Tt);
            for (int tt = 0; t < t_end; ++t, ++tt) {
                const double *qbf_cur = qbf_s + (t & 1) * qbf_stride;
                double *qbf_nxt = qbf_s + ((t & 1) ^ 1) * qbf_stride;
                if (is_compute) {
                    if (ia >= 0) {
                        const double p = rain_t[tt * C.ngp + h.ippt], ep = pet_t[tt * C.nge + h.ipet];
                        // ---- dynamic_dist gather (dyna_dynamic_dist.f90:49-52) ----
                        double qin[S];
#pragma unroll
                        for (int k = 0; k < S; ++k) qin[k] = 0.0;
                        for (int e = h.row_beg; e < h.row_end; ++e) {
                            const int src = src_s[e];
                            const double w = w_s[e];
                            const double2 *qv = reinterpret_cast<const double2 *>(qbf_cur + (gbase + src) * S);
                            double qs[S];
#pragma unroll
                            for (int k2 = 0; k2 < S / 2; ++k2) { const double2 v = qv[k2]; qs[2 * k2] = v.x; qs[2 * k2 + 1] = v.y; }
                            if (FAST) {
#pragma unroll
                                for (int k = 0; k < S; ++k) qin[k] = qin[k] + qs[k] * w;
                            } else {
                                const double acs = ac_s[src];
#pragma unroll
                                for (int k = 0; k < S; ++k) qin[k] = qin[k] + qs[k] * w * acs;
                            }
                        }
                        // ---- root zone, unsaturated zone, saturated zone per member slot ----
                        double *qo = qbf_nxt + (gbase + ia) * S;
                        double *exo = ex_s + (t & 1) * ex_stride + (g * S) * nac_pad + ia;
                        const double *cst = cst_s + (gbase + ia) * S;
                        if (LEAN) {
                            const double epp = fmax(ep, 0.0);
#pragma unroll
                            for (int k0 = 0; k0 < S; k0 += DCP_RES_LG) {
                                LeanOut lo[DCP_RES_LG];
                                double qi_old[DCP_RES_LG];
#pragma unroll
                                for (int j = 0; j < DCP_RES_LG; ++j) {
                                    const int k = k0 + j;
                                    const double *pp = par_s + ((size_t)(g * S + k) * npt + h.ipar) * RES_PARW;
                                    HruConst c;
                                    c.ac = h.ac; c.inv_ac = inv_ac;
                                    c.q0 = cst[k]; c.q2 = cst[qbf_stride + k]; c.m2 = cst[2 * qbf_stride + k];
                                    c.q1 = cst[3 * qbf_stride + k]; c.inv_m2 = cst[4 * qbf_stride + k];
                                    c.srmax = pp[0]; c.td = pp[1]; c.smax = pp[2]; c.inv_srmax = pp[4];
                                    c.ims_rz = 1; c.ims_uz = 1;
                                    qi_old[j] = qint1[k];
                                    HruState st;
                                    st.sd = sd[k]; st.suz = suz[k]; st.srz = srz[k]; st.qint1 = qint1[k];
                                    hru_step_lean(c, gc, p, epp, qin[k], st, lo[j]);
                                    sd[k] = st.sd; suz[k] = st.suz; srz[k] = st.srz; qint1[k] = st.qint1;
                                }
                                bool any_slow = false;
#pragma unroll
                                for (int j = 0; j < DCP_RES_LG; ++j) any_slow |= lo[j].slow;
                                if (any_slow) {
#pragma unroll
                                    for (int j = 0; j < DCP_RES_LG; ++j)
                                        if (lo[j].slow) {           // rare: constants are re-read instead of kept live
                                            const int k = k0 + j;
                                            const double *pp = par_s + ((size_t)(g * S + k) * npt + h.ipar) * RES_PARW;
                                            HruConst c;
                                            c.ac = h.ac; c.inv_ac = inv_ac;
                                            c.q0 = cst[k]; c.q2 = cst[qbf_stride + k]; c.m2 = cst[2 * qbf_stride + k];
                                            c.q1 = cst[3 * qbf_stride + k]; c.inv_m2 = cst[4 * qbf_stride + k];
                                            c.smax = pp[2];
                                            hru_step_lean_fixup(c, gc, qin[k], qi_old[j], sd[k], lo[j]);
                                        }
                                }
#pragma unroll
                                for (int j = 0; j < DCP_RES_LG; ++j) {
                                    qo[k0 + j] = lo[j].qbf;
                                    exo[(k0 + j) * nac_pad] = lo[j].exac;
                                    if (AET) aet[k0 + j] = aet[k0 + j] + lo[j].ea;
                                }
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < S; ++k) {
                                const double *pp = par_s + ((size_t)(g * S + k) * npt + h.ipar) * RES_PARW;
                                HruConst c;
                                c.ac = h.ac; c.inv_ac = inv_ac;
                                c.q0 = cst[k]; c.q1 = cst[3 * qbf_stride + k]; c.q2 = cst[qbf_stride + k];
                                c.srmax = pp[0]; c.td = pp[1]; c.smax = pp[2];
                                c.inv_srmax = pp[4];
                                c.m2 = cst[2 * qbf_stride + k];
                                c.inv_m2 = cst[4 * qbf_stride + k];
                                c.ims_rz = h.ims_rz; c.ims_uz = h.ims_uz;
                                HruState s;
                                s.sd = sd[k]; s.suz = suz[k]; s.srz = srz[k]; s.qint1 = qint1[k];
                                HruFlux f;
                                hru_step<FAST, false>(c, gc, p, ep, qin[k], s, f);
                                sd[k] = s.sd; suz[k] = s.suz; srz[k] = s.srz; qint1[k] = s.qint1;
                                qo[k] = f.qbf;
                                exo[k * nac_pad] = (f.ex > 0.0) ? f.ex * h.ac : 0.0;   // dyna_results_ac.f90:35-38
                                if (AET) { if (h.ims_rz == 1 && ep > 0.0) aet[k] = aet[k] + f.ea; }
                            }
                        }
                    }
                } else {
                    // ---- river warp: qb(t) from qbf_cur, qof(t-1) from last step's ex terms ----
                    const int lane = tid - RES_COMPUTE;
                    const double *ex_prev = ex_s + ((t & 1) ^ 1) * ex_stride;
#pragma unroll
                    for (int q = 0; q < RES_MAXPASS; ++q) {
                        const int cidx = lane + 32 * q;
                        if (cidx < Mc * nriv) {
                            const int jm = cidx / nriv, r = cidx - jm * nriv;
                            const int gg = jm / S, kk = jm - gg * S;
                            const int m = m_base + jm;
                            double qb = 0.0;
                            for (int i = P.rivq_beg[r]; i < P.rivq_beg[r + 1]; ++i) {
                                const RivChain rc = P.rivq[i];
                                const double qv = qbf_cur[(gg * nac_pad + rc.ia) * S + kk];
                                qb = qb + C.dtt * qv * rc.w_riv * rc.share * rc.ac;    // dyna_dynamic_dist.f90:59-61
                            }
                            if (t > 0) {
                                double qof = 0.0;
                                for (int i = P.rivo_beg[r]; i < P.rivo_beg[r + 1]; ++i)
                                    qof = qof + ex_prev[jm * nac_pad + P.rivo[i]];
                                if (m < B) W.qout[((size_t)((t - 1) & Rm) * nriv + r) * B + m] = qb_prev[q] + qof;   // dyna_river.f90:43
                            }
                            qb_prev[q] = qb;
                        }
                    }
                }
                __syncthreads();
            }
        }
        // ---- drain: qof of the last step ----
        if (!is_compute) {
            const int lane = tid - RES_COMPUTE;
            const double *ex_prev = ex_s + ((t & 1) ^ 1) * ex_stride;
#pragma unroll
            for (int q = 0; q < RES_MAXPASS; ++q) {
                const int cidx = lane + 32 * q;
                if (cidx < Mc * nriv) {
                    const int jm = cidx / nriv, r = cidx - jm * nriv;
                    const int m = m_base + jm;
                    double qof = 0.0;
                    for (int i = P.rivo_beg[r]; i < P.rivo_beg[r + 1]; ++i) qof = qof + ex_prev[jm * nac_pad + P.rivo[i]];
                    if (m < B) W.qout[((size_t)((t - 1) & Rm) * nriv + r) * B + m] = qb_prev[q] + qof;
                }
            }
        } else if (AET) {
#pragma unroll
            for (int k = 0; k < S; ++k)
                if (ia >= 0 && m_base + g * S + k < B) W.aet[(size_t)ia * B + (m_base + g * S + k)] = aet[k];
        }
    }
}

// ---------------------------------------------------------------------------------------------
template <int S>
static cudaError_t launch_res(const DevCatchment &C, const DevBatch &W, const ResPlan &P, bool fast, bool aet,
                              int grid, cudaStream_t s) {
    void (*k)(DevCatchment, DevBatch, ResPlan);
    const int mode = !fast ? 0 : (P.lean_ok ? 2 : 1);
    if (mode == 2) k = aet ? k_physics_resident<S, 2, true> : k_physics_resident<S, 2, false>;
    else if (mode == 1) k = aet ? k_physics_resident<S, 1, true> : k_physics_resident<S, 1, false>;
    else k = aet ? k_physics_resident<S, 0, true> : k_physics_resident<S, 0, false>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, P.smem_bytes);
    if (e != cudaSuccess) return e;
    k<<<grid, RES_THREADS, P.smem_bytes, s>>>(C, W, P);
    return cudaGetLastError();
}

cudaError_t dcp_launch_physics_resident(const DevCatchment &C, const DevBatch &W, const ResPlan &P, bool fast,
                                        bool aet, int grid, cudaStream_t s) {
    return launch_res<DCP_RES_S>(C, W, P, fast, aet, grid, s);
}
