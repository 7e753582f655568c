// dcp_stepwise.cu — the Topmod time loop for catchments that do not fit on chip (nac > 512) and/or
// have a wide dynamic_dist weighting matrix (BASELINE config 4): one timestep = three launches,
//
//   (1) redistribute   QIN[nac x B] = What[nac x nac] . QBF_old[nac x B]      (dyna_dynamic_dist.f90:47-52)
//         sparse W : k_redistribute_csr  — destination-major CSR, thread = member, grid.y = row chunks;
//                    sources ascending, (qbf*w)*ac per term: the reference's scatter order bit for bit
//         dense  W : k_redistribute_dense — a register-tiled FP64 GEMM (128x128x16 block tile, 8x8 per
//                    thread, double-buffered shared memory, all-zero 128x16 tiles of W skipped); every
//                    output is accumulated by ONE thread over ascending source index, so EXACT mode
//                    ((qbf*w)*ac, separate multiply/add) still reproduces the reference's order and FAST
//                    mode uses one FMA per term.  Chosen when nnz(W)/nac^2 >= DCP_DENSE_MIN_DENSITY.
//   (2) k_physics_step  root zone -> unsaturated zone -> saturated zone for every (HRU, member) of the step,
//                    state streamed through HBM member-fastest (80 B/unit); the HRUs of a step are independent
//                    once QIN is known, so the grid covers (members, HRU chunks) instead of members only;
//                    each HRU leaves its terms of qb(r) and qof(r) in HBM
//   (3) k_river_sum   qout(r) = qb(r) + qof(r), both summed over ascending HRU index by one thread per
//                    (member, river): the reference's order (dyna_dynamic_dist.f90:57-66, dyna_results_ac.f90:35-38,
//                    dyna_river.f90:43).
//
// Why not tensor cores for (1): B200's FP64 MMA rate equals its DFMA rate (both 64 FMA/clk/SM), the GEMM below is
// FP64-pipe bound with 8 LDS.128 per 64 DFMA, and a single accumulation chain per output keeps the reference's
// summation order.  Built with -fmad=false like the other exact kernels; FAST contractions are explicit fma().
#include "dcp_device.cuh"
#include "dcp_kernels.h"
#include "dcp_physics.cuh"

#include <cstdint>
#include <cstdlib>

#define SW_THREADS 128

// ------------------------------------------------------------------------------------------------
// (1a) sparse redistribute: thread = member, blockIdx.y = chunk of destination rows
// ------------------------------------------------------------------------------------------------
template <bool ONLY_FLAGGED>
__global__ void __launch_bounds__(SW_THREADS)
k_redistribute_csr(DevCatchment C, DevBatch W, const double *__restrict__ qbf_old, int rows_per_chunk) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    const int B = W.B;
    if (m >= B) return;
    if (ONLY_FLAGGED && W.bad_src[m] == 0) return;
    const int ia0 = blockIdx.y * rows_per_chunk, ia1 = min(C.nac, ia0 + rows_per_chunk);
    for (int ia = ia0; ia < ia1; ++ia) {
        const int rb = C.hru[ia].row_beg, re = C.hru[ia].row_end;
        double qin = 0.0;
        for (int e = rb; e < re; ++e) {
            const WEntry we = C.w[e];
            qin = qin + qbf_old[we.src * (size_t)B + m] * we.w * we.ac_src;
        }
        W.qin_ext[ia * (size_t)B + m] = qin;
    }
}

// ------------------------------------------------------------------------------------------------
// (1b) dense redistribute: C[Mp x B] = A[Mp x Kp] . Bm[Kp x B], A = W in destination-major dense form
// ------------------------------------------------------------------------------------------------
#define GM_BM 128
#define GM_BN 128
#define GM_BK 16
#define GM_THREADS 256

// A non-finite baseflow must reach its true targets only (0 * inf would poison every destination of the member):
// the GEMM stages 0 instead and flags the member, whose rows are then redone by k_redistribute_csr<true> in
// the reference's sparse form.
__device__ __forceinline__ bool is_finite_bits(double x) { return (__double2hiint(x) & 0x7ff00000) != 0x7ff00000; }

template <bool EXACT>
__global__ void __launch_bounds__(GM_THREADS, 1)
k_redistribute_dense(DensePlan D, const double *__restrict__ Bm, double *__restrict__ Cm, int32_t *__restrict__ bad, int nac, int B) {
    extern __shared__ __align__(16) double gsm[];
    double *As = gsm;                                     // [2][BK][BM]   (k-major: transposed W tile)
    double *Bs = gsm + 2 * GM_BK * GM_BM;                 // [2][BK][BN]
    double *Cs = Bs + 2 * GM_BK * GM_BN;                  // [2][BK]       ac(src) (EXACT)
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n0 = blockIdx.x * GM_BN, bm = D.bm_order[blockIdx.y], m0 = bm * GM_BM;
    const int kt_b = D.kt_beg[bm], kt_e = D.kt_beg[bm + 1];

    // global -> register staging: A: row tid/2, k half (tid&1)*8;  B: k row tid/16, members (tid&15)*8
    const int a_r = tid >> 1, a_k = (tid & 1) * 8;
    const int b_k = tid >> 4, b_n = (tid & 15) * 8;
    double ra[8], rb[8], rc = 0.0, rsc = 0.0;
    // gload only ISSUES the global loads (the values are first touched in sstore, after the math of the current
    // tile), so their latency hides behind 1024 DFMAs per thread
    auto gload = [&](int kt) {
        const int k0 = D.kt_idx[kt] * GM_BK;
        const double2 *ap = reinterpret_cast<const double2 *>(D.wd + (size_t)(m0 + a_r) * D.Kp + k0 + a_k);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const double2 v = __ldg(ap + i); ra[2 * i] = v.x; ra[2 * i + 1] = v.y; }
        const int k = k0 + b_k;
        rsc = (k < nac) ? D.acs[k] : 0.0;
        const double *bp = Bm + (size_t)k * B + n0 + b_n;
        if (k < nac && n0 + b_n + 8 <= B && (B & 1) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { const double2 v = reinterpret_cast<const double2 *>(bp)[i]; rb[2 * i] = v.x; rb[2 * i + 1] = v.y; }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) rb[i] = (k < nac && n0 + b_n + i < B) ? bp[i] : 0.0;
        }
        if (tid < GM_BK) { const int kk = k0 + tid; rc = kk < nac ? D.acs[kk] : 0.0; }
    };
    auto sstore = [&](int buf) {
        double *a = As + buf * GM_BK * GM_BM, *b = Bs + buf * GM_BK * GM_BN;
#pragma unroll
        for (int i = 0; i < 8; ++i) a[(a_k + i) * GM_BM + a_r] = ra[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (!is_finite_bits(rb[i])) { bad[n0 + b_n + i] = 1; rb[i] = 0.0; }
            if (!EXACT) rb[i] = rb[i] * rsc;              // FAST: ac(src) folded into the staged tile
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) reinterpret_cast<double2 *>(b + b_k * GM_BN + b_n)[i] = make_double2(rb[2 * i], rb[2 * i + 1]);
        if (tid < GM_BK) Cs[buf * GM_BK + tid] = rc;
    };

    double acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;

    if (kt_b < kt_e) {
        gload(kt_b);
        sstore(0);
        __syncthreads();
        for (int kt = kt_b; kt < kt_e; ++kt) {
            const int buf = (kt - kt_b) & 1;
            if (kt + 1 < kt_e) gload(kt + 1);             // next tile's global loads fly during this tile's math
            const double *a = As + buf * GM_BK * GM_BM + ty * 8;
            const double *b = Bs + buf * GM_BK * GM_BN + tx * 2;
#pragma unroll
            for (int k = 0; k < GM_BK; ++k) {
                double av[8], bv[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double2 v = reinterpret_cast<const double2 *>(a + k * GM_BM)[i];
                    av[2 * i] = v.x; av[2 * i + 1] = v.y;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double2 v = *reinterpret_cast<const double2 *>(b + k * GM_BN + j * 32);
                    bv[2 * j] = v.x; bv[2 * j + 1] = v.y;
                }
                if (EXACT) {
                    const double ac = Cs[buf * GM_BK + k];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[i][j] = acc[i][j] + (bv[j] * av[i]) * ac;     // (qbf*w)*ac, dyna_dynamic_dist.f90:50-51
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
                }
            }
            if (kt + 1 < kt_e) sstore(buf ^ 1);
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = m0 + ty * 8 + i;
        if (r >= nac) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + j * 32 + tx * 2;
            double *cp = Cm + (size_t)r * B + n;
            if (n + 1 < B && (B & 1) == 0) *reinterpret_cast<double2 *>(cp) = make_double2(acc[i][2 * j], acc[i][2 * j + 1]);
            else { if (n < B) cp[0] = acc[i][2 * j]; if (n + 1 < B) cp[1] = acc[i][2 * j + 1]; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// (1c) dense redistribute on the FP64 tensor cores (FAST policy): the same 128 x 128 x 16 block tile and staging as (1b),
// the inner product as mma.sync.m8n8k4.f64 (SASS DMMA).  8 warps as 2 (destinations) x 4 (members), warp tile 64 x 32 =
// 8 x 4 MMA tiles, 32 accumulator pairs per lane.  Per k-step of 4 a warp issues 12 LDS.64 and 32 DMMA (8 192 FMAs) where
// the vector kernel issues 32 LDS.128 and 256 DFMA: B200's FP64 tensor rate equals its DFMA rate, so the gain is issue
// slots and operand bandwidth, not peak (A/B with ncu: profiles/r2_c4_dense_mma_vs_dfma.txt).
// Shared tiles are k-major with a row stride of 136 doubles: the 4 k-rows x 8 consecutive columns of a fragment load then
// fall into 32 distinct bank pairs (stride = 8 mod 16 doubles).
// The accumulation order inside an MMA is the hardware's, so this kernel exists for the FAST policy only.
// ------------------------------------------------------------------------------------------------
#define GM_LD 136
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(GM_THREADS, 1)
k_redistribute_dense_mma(DensePlan D, const double *__restrict__ Bm, double *__restrict__ Cm, int32_t *__restrict__ bad, int nac, int B) {
    extern __shared__ __align__(16) double gsm[];
    double *As = gsm;                                     // [2][BK][GM_LD]   (k-major: transposed W tile)
    double *Bs = gsm + 2 * GM_BK * GM_LD;                 // [2][BK][GM_LD]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;        // warp tile origin inside the block tile
    const int fr = lane >> 2, fk = lane & 3;                      // fragment row / k index of this lane
    const int n0 = blockIdx.x * GM_BN, bm = D.bm_order[blockIdx.y], m0 = bm * GM_BM;
    const int kt_b = D.kt_beg[bm], kt_e = D.kt_beg[bm + 1];
    const int a_r = tid >> 1, a_k = (tid & 1) * 8;
    const int b_k = tid >> 4, b_n = (tid & 15) * 8;
    double ra[8], rb[8], rsc = 0.0;
    auto gload = [&](int kt) {
        const int k0 = D.kt_idx[kt] * GM_BK;
        const double2 *ap = reinterpret_cast<const double2 *>(D.wd + (size_t)(m0 + a_r) * D.Kp + k0 + a_k);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const double2 v = __ldg(ap + i); ra[2 * i] = v.x; ra[2 * i + 1] = v.y; }
        const int k = k0 + b_k;
        rsc = (k < nac) ? D.acs[k] : 0.0;
        const double *bp = Bm + (size_t)k * B + n0 + b_n;
        if (k < nac && n0 + b_n + 8 <= B && (B & 1) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { const double2 v = reinterpret_cast<const double2 *>(bp)[i]; rb[2 * i] = v.x; rb[2 * i + 1] = v.y; }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) rb[i] = (k < nac && n0 + b_n + i < B) ? bp[i] : 0.0;
        }
    };
    auto sstore = [&](int buf) {
        double *a = As + buf * GM_BK * GM_LD, *b = Bs + buf * GM_BK * GM_LD;
#pragma unroll
        for (int i = 0; i < 8; ++i) a[(a_k + i) * GM_LD + a_r] = ra[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (!is_finite_bits(rb[i])) { bad[n0 + b_n + i] = 1; rb[i] = 0.0; }
            rb[i] = rb[i] * rsc;                          // ac(src) folded into the staged tile
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) reinterpret_cast<double2 *>(b + b_k * GM_LD + b_n)[i] = make_double2(rb[2 * i], rb[2 * i + 1]);
    };

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    if (kt_b < kt_e) {
        gload(kt_b);
        sstore(0);
        __syncthreads();
        for (int kt = kt_b; kt < kt_e; ++kt) {
            const int buf = (kt - kt_b) & 1;
            if (kt + 1 < kt_e) gload(kt + 1);
            const double *a = As + buf * GM_BK * GM_LD + wm + fr;
            const double *b = Bs + buf * GM_BK * GM_LD + wn + fr;
#pragma unroll
            for (int k4 = 0; k4 < GM_BK; k4 += 4) {
                double af[8], bf[4];
#pragma unroll
                for (int i = 0; i < 8; ++i) af[i] = a[(k4 + fk) * GM_LD + i * 8];
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = b[(k4 + fk) * GM_LD + j * 8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
            if (kt + 1 < kt_e) sstore(buf ^ 1);
            __syncthreads();
        }
    }
    // C fragment: row = lane / 4, columns 2 * (lane % 4) + {0, 1} of each 8 x 8 tile
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = m0 + wm + i * 8 + fr;
        if (r >= nac) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + wn + j * 8 + 2 * fk;
            double *cp = Cm + (size_t)r * B + n;
            if (n + 1 < B && (B & 1) == 0) *reinterpret_cast<double2 *>(cp) = make_double2(acc[i][j][0], acc[i][j][1]);
            else { if (n < B) cp[0] = acc[i][j][0]; if (n + 1 < B) cp[1] = acc[i][j][1]; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// (2) one timestep of every (HRU, member): thread = member, blockIdx.y = HRU chunk
// ------------------------------------------------------------------------------------------------
template <bool FAST, bool AET, bool NPT1>
__global__ void __launch_bounds__(SW_THREADS)
k_physics_step(DevCatchment C, DevBatch W, const int *__restrict__ t_base, int t_off, int rows_per_chunk) {
    const int t = *t_base + t_off;                 // the launch set of a step is replayed from a CUDA graph: the timestep
                                                   // comes from a device counter (+ the node's offset inside the graph)
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    const int B = W.B;
    if (m >= B) return;
    const int npt = C.npt;
    const int ia0 = blockIdx.y * rows_per_chunk, ia1 = min(C.nac, ia0 + rows_per_chunk);
    StepConst g;
    g.dt = C.dt; g.dtt = C.dtt; g.inv_dtt = C.inv_dtt; g.ntt = C.ntt;
    double p_szm = 0, p_srmax = 0, p_td = 0, p_smax = 0;
    if (NPT1) {
        p_szm = W.par[(0 * npt) * (size_t)B + m];
        p_srmax = W.par[(2 * npt) * (size_t)B + m];
        p_td = W.par[(5 * npt) * (size_t)B + m];
        p_smax = W.par[(6 * npt) * (size_t)B + m];
    }
    const double inv_szm = 1.0 / p_szm, inv_srmax = 1.0 / p_srmax;
    const double *__restrict__ qbf_old = (t & 1) ? W.qbf1 : W.qbf0;
    double *__restrict__ qbf_new = (t & 1) ? W.qbf0 : W.qbf1;
    const double *rain_t = C.rain + (size_t)t * C.ngp;
    const double *pet_t = C.pet + (size_t)t * C.nge;

    for (int ia = ia0; ia < ia1; ++ia) {
        const HruStatic h = C.hru[ia];
        const size_t o = ia * (size_t)B + m;
        // dynamic_dist, river part (dyna_dynamic_dist.f90:57-66): this HRU's terms (((dtt*qbf)*w_riv)*rivers)*ac
        {
            const double base = C.dtt * qbf_old[o] * h.w_riv;
            for (int e = h.riv_beg; e < h.riv_end; ++e) {
                const RivEntry re = C.riv[e];
                W.qb_term[e * (size_t)B + m] = base * re.share * h.ac;
            }
        }
        const double qin = W.qin_ext[o];
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
        const double p = __ldg(rain_t + h.ippt), ep = __ldg(pet_t + h.ipet);
        HruFlux f;
        hru_step<FAST, false>(c, g, p, ep, qin, s, f);
        W.sd[o] = s.sd; W.suz[o] = s.suz; W.srz[o] = s.srz; W.qint1[o] = s.qint1;
        qbf_new[o] = f.qbf;
        if (AET) { if (h.ims_rz == 1 && ep > 0.0) W.aet[o] = W.aet[o] + f.ea; }
        W.qof_term[o] = (f.ex > 0.0) ? f.ex * h.ac : 0.0;                  // dyna_results_ac.f90:35-38
    }
    if (blockIdx.y == 0) W.bad_src[m] = 0;          // the redistribute of the next step flags afresh
}

// ------------------------------------------------------------------------------------------------
// (3) river sums in the reference's order: thread = member, blockIdx.y = river
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SW_THREADS)
k_river_sum(DevCatchment C, DevBatch W, DensePlan D, const int *__restrict__ t_base, int t_off) {
    const int t = *t_base + t_off;
    const int m = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    const int B = W.B;
    if (m >= B) return;
    // eight loads in flight per batch, added in list order (the loads are independent, the sum is not)
    auto ordered_sum = [&](const double *__restrict__ term, const int32_t *__restrict__ idx, int beg, int end) {
        double acc = 0.0;
        int i = beg;
        for (; i + 8 <= end; i += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = term[__ldg(idx + i + u) * (size_t)B + m];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = acc + v[u];
        }
        for (; i < end; ++i) acc = acc + term[__ldg(idx + i) * (size_t)B + m];
        return acc;
    };
    const double qb = ordered_sum(W.qb_term, D.rq_idx, D.rq_beg[r], D.rq_beg[r + 1]);
    const double qof = ordered_sum(W.qof_term, D.ro_idx, D.ro_beg[r], D.ro_beg[r + 1]);
    W.qout[((size_t)m * C.nriv + r) * W.R + (t & W.Rmask)] = qb + qof;      // dyna_river.f90:43
}

// ------------------------------------------------------------------------------------------------
// launch wrapper: steps [t0, t1)
// ------------------------------------------------------------------------------------------------
__global__ void k_step_counter(int *t_base, int set_to, int add) { *t_base = set_to >= 0 ? set_to : *t_base + add; }

template <bool FAST, bool AET>
static void launch_step(const DevCatchment &C, const DevBatch &W, const int *t_base, int t_off, int rows, dim3 grid, cudaStream_t s) {
    if (C.npt == 1) k_physics_step<FAST, AET, true><<<grid, SW_THREADS, 0, s>>>(C, W, t_base, t_off, rows);
    else k_physics_step<FAST, AET, false><<<grid, SW_THREADS, 0, s>>>(C, W, t_base, t_off, rows);
}

#define SW_GRAPH_STEPS 32             // timesteps per CUDA graph (even: the baseflow double buffer repeats with period 2)

cudaError_t dcp_launch_physics_stepwise(const DevCatchment &C, const DevBatch &W, const DensePlan &D, int t0, int t1,
                                        bool fast, bool aet, bool dense, int n_sm, cudaStream_t s) {
    const int gx = (W.B + SW_THREADS - 1) / SW_THREADS;
    // HRU chunks: enough thread blocks to fill the GPU a few times over
    int n_chunk = (n_sm * 12 + gx - 1) / gx;
    n_chunk = max(1, min(n_chunk, (C.nac + 15) / 16));
    const int rows = (C.nac + n_chunk - 1) / n_chunk;
    n_chunk = (C.nac + rows - 1) / rows;
    const dim3 g_phys(gx, n_chunk), g_riv(gx, C.nriv);
    const size_t gsm = (size_t)(2 * GM_BK * GM_BM + 2 * GM_BK * GM_BN + 2 * GM_BK) * sizeof(double);
    if (dense) {
        cudaError_t e = cudaFuncSetAttribute(fast ? k_redistribute_dense<false> : k_redistribute_dense<true>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm);
        if (e != cudaSuccess) return e;
    }
    const dim3 g_gemm((W.B + GM_BN - 1) / GM_BN, D.Mp / GM_BM);
    // FAST policy: the tensor-core contraction unless DCP_DENSE_MMA=0 asks for the vector kernel (A/B runs)
    const size_t gsm_mma = (size_t)(4 * GM_BK * GM_LD) * sizeof(double);
    const bool use_mma = dense && fast && !(getenv("DCP_DENSE_MMA") && atoi(getenv("DCP_DENSE_MMA")) == 0);
    if (use_mma) {
        cudaError_t e = cudaFuncSetAttribute(k_redistribute_dense_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm_mma);
        if (e != cudaSuccess) return e;
    }
    // one timestep = 3-4 kernels; the launch set of SW_GRAPH_STEPS consecutive steps is captured ONCE into a CUDA graph and
    // replayed, so host launches do not scale with nstep (8 760 steps = 274 graph launches instead of 35 040 kernel launches).
    // Kernel arguments are frozen in a graph: the timestep is read from a device counter that the graph's last node advances.
    int *t_base = W.t_dev;
    auto enqueue_step = [&](int t_abs, int t_off) {         // t_abs only selects the baseflow buffer (parity)
        const double *qbf_old = (t_abs & 1) ? W.qbf1 : W.qbf0;
        if (dense) {
            if (use_mma) k_redistribute_dense_mma<<<g_gemm, GM_THREADS, gsm_mma, s>>>(D, qbf_old, W.qin_ext, W.bad_src, C.nac, W.B);
            else if (fast) k_redistribute_dense<false><<<g_gemm, GM_THREADS, gsm, s>>>(D, qbf_old, W.qin_ext, W.bad_src, C.nac, W.B);
            else k_redistribute_dense<true><<<g_gemm, GM_THREADS, gsm, s>>>(D, qbf_old, W.qin_ext, W.bad_src, C.nac, W.B);
            k_redistribute_csr<true><<<g_phys, SW_THREADS, 0, s>>>(C, W, qbf_old, rows);      // members flagged non-finite only
        } else {
            k_redistribute_csr<false><<<g_phys, SW_THREADS, 0, s>>>(C, W, qbf_old, rows);
        }
        if (aet) { if (fast) launch_step<true, true>(C, W, t_base, t_off, rows, g_phys, s); else launch_step<false, true>(C, W, t_base, t_off, rows, g_phys, s); }
        else { if (fast) launch_step<true, false>(C, W, t_base, t_off, rows, g_phys, s); else launch_step<false, false>(C, W, t_base, t_off, rows, g_phys, s); }
        k_river_sum<<<g_riv, SW_THREADS, 0, s>>>(C, W, D, t_base, t_off);
    };
    int t = t0;
    k_step_counter<<<1, 1, 0, s>>>(t_base, t0, 0);
    if (t & 1) { enqueue_step(t, 0); k_step_counter<<<1, 1, 0, s>>>(t_base, -1, 1); ++t; }      // graphs start on an even step
    const int n_graph = (t1 - t) / SW_GRAPH_STEPS;
    if (n_graph >= 2 && !getenv("DCP_STEPWISE_NO_GRAPH")) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) return e;
        for (int i = 0; i < SW_GRAPH_STEPS; ++i) enqueue_step(t + i, i);
        k_step_counter<<<1, 1, 0, s>>>(t_base, -1, SW_GRAPH_STEPS);
        e = cudaStreamEndCapture(s, &graph);
        if (e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
        if (e != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return e; }
        for (int g2 = 0; g2 < n_graph && e == cudaSuccess; ++g2) e = cudaGraphLaunch(exec, s);
        t += n_graph * SW_GRAPH_STEPS;
        // the executable graph must outlive its launches: the stream is synchronised before it is destroyed
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        cudaGraphExecDestroy(exec);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return e;
    }
    for (int i = 0; t + i < t1; ++i) enqueue_step(t + i, i);      // remainder (and short tiles): direct launches off the same counter
    return cudaGetLastError();
}
