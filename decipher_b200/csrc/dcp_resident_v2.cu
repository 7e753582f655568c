// dcp_resident_v2.cu — second-generation persistent on-chip kernel for the lean (branch-free, fast
// arithmetic) step.  Same residency model as dcp_resident.cu (a member's stores stay in registers and
// its last baseflow in shared memory for the WHOLE simulation; rain/PET tiles arrive by TMA), rebuilt
// around what the ncu captures of the first generation showed (profiles/r1_ncu_resident_*, r1_ncu_v2_*):
//
//   (1) the dynamic_dist gather is bound by shared-memory BANDWIDTH, not latency: every lane reads a
//       different source tuple, 32-byte tuples collide on four bank classes, and all warps of the CTA
//       gather at the same time (right after the per-step barrier) while the FP64 pipe idles;
//   (2) the river warps were as busy as the compute warps: three loads per summed term;
//   (3) one warp owned the longest W rows and every other warp waited for it at the barrier.
//
// Design:
//   * TWO independent teams per CTA ("virtual CTAs"): team = 4 compute warps + 2 river warps with its
//     own named barrier, forcing pipeline and member groups; a team runs 2 member slots over all 512 HRU
//     columns (4 columns per thread).  Members never interact, so the teams drift out of phase and one
//     team's gather (shared-memory bound) overlaps the other's physics (FP64 bound).  Only the weight
//     matrix and the per-column tables are shared (read-only).
//   * gather = ELLPACK SpMM over one flat list per thread.  W is stored per block of 32 thread positions,
//     step-major, rows padded with zero weights to the block's longest row; the four columns of a thread follow
//     one another in the list, so the trip counts are warp-uniform and the list is walked with a one-entry
//     look-ahead (the first entry of the next step is requested before the barrier).  An entry is ONE 8-byte
//     word: the weight with the source tuple's byte offset (index << 4) in mantissa bits 4..12 (|dw/w| <= 2^-40), i.e.
//     one LDS.64 per entry, no index array, and the tuple address is one LOP3 (the tuple buffers are 8 KB aligned).  A tuple is the 16 bytes of the team's two member slots, so one
//     LDS.128 per source and eight bank classes; the host orders the entries of each row so that the eight
//     lanes of a quarter-warp hit different classes wherever the catchment allows (dcp_api.cu).
//     Measured (profiles/r1_ncu_resident_v2*.txt): the gather is ~28 % of a compute warp's time for ~14 % of
//     its instructions (LDS latency + MIO queue), the reason the FP64 pipe sits near 40 %.
//   * exp(sd/m2) is carried in the state instead of being recomputed (hru_step_lean<true>, dcp_physics.cuh).
//   * river sums by contribution.  Every compute thread publishes, per member slot,
//         c = (dtt*w_riv*share*ac) * qbf_old + (exs+exus)*ac          (dyna_dynamic_dist.f90:59-61,
//                                                                      dyna_results_ac.f90:35-38)
//     into a river-major slot array, so qout(t) = qb(t) + qof(t) (dyna_river.f90:43) is the plain sum of a
//     contiguous range: the river warps read tuples with no index or coefficient loads and finish with a
//     shuffle tree.  Needs every HRU's river share to sit on its own river `ipriv` (true for DTA output:
//     an HRU class carries its gauge id); other catchments use the first generation.
//   * blocks of 32 positions are dealt to the four compute warps longest-first (LPT) so that the
//     per-step gather work of the warps is even.
//   * (round 2) river ROUTING and the OBJECTIVE sums inside the kernel: the river warps store qout(t) into a ring per
//     member in flight, and once per 64-step forcing tile the team's compute warps route the finished tile with lane =
//     timestep (r2_route_tile: route_process_run_step, DTA/dta_route_processing.f90:671-727, + dyna_river.f90:57-59);
//     only measures leave the chip.  Measured alternatives (routing warp, routing in the river warps, a split per-step
//     barrier): profiles/r2_ab_routing_designs.txt.
//   * (round 2) table-driven exp / log in the lean step (fm_exp_tab / fm_log_tab, tables in shared memory): 84 instead
//     of 102 executed FP64 instructions per unit.
//   * (round 2) a second build of this file, dcp_resident_v2w.cu, holds 1024 columns per member slot (513 .. 1024 HRUs).
//
// Built WITH FMA contraction (no operation-order contract in the lean policy; parity is checked against
// the oracle at 1e-9).  Hot path citations: RRMODEL/dyna_topmod.f90:92-233, dyna_dynamic_dist.f90:47-67,
// dyna_root_zone1.f90, dyna_unsat_zone1.f90, dyna_satzone_analyticalsolver.f90, dyna_river.f90:42-49.
#include "dcp_device.cuh"
#include "dcp_physics.cuh"
#include "dcp_kernels.h"
#include "dcp_resident_common.cuh"

#include <cstdint>

// Two builds of this file (the second through dcp_resident_v2w.cu, which defines R2_WIDE and includes it):
//   narrow  2 teams x (4 compute + 2 river warps), 512 unit columns per member slot   -- catchments of up to 512 HRUs
//   WIDE    1 team  x (8 compute + 4 river warps), 1024 unit columns per member slot  -- 513 .. 1024 HRUs
// Same thread count, register split and code; the wide build trims the per-column tables to fit 1024 columns in shared
// memory: 1/ac is folded into the gather weights, q0 = exp(t0dt) * exp(-st) is rebuilt from a per-member and a per-column
// factor instead of being stored per unit, and the column words shrink from three to two (+ the exp(-st) factor).
#ifndef R2_WIDE
#define R2_WIDE 0
#endif
#if R2_WIDE
#define R2_TEAMS 1
#define R2_KERNEL k_physics_resident_v2w
#define R2_LAUNCH dcp_launch_physics_resident_v2w
#else
#define R2_TEAMS 2
#define R2_KERNEL k_physics_resident_v2
#define R2_LAUNCH dcp_launch_physics_resident_v2
#endif
#define R2_TC (256 / R2_TEAMS)        // compute threads per team (4 columns each)
#define R2_COLS (4 * R2_TC)           // unit columns per member slot: 512 / 1024
#define R2_TR (128 / R2_TEAMS)        // river threads per team: contribution sums one step behind the compute warps; lane 0 also
                                      // drives the forcing TMA
#define R2_TT (R2_TC + R2_TR)         // threads at a team's per-step barrier
#define R2_CT (R2_TEAMS * R2_TC)      // 256 compute threads
#define R2_THREADS (R2_TEAMS * (R2_TC + R2_TR)) // 384
#define R2_TUPB (R2_COLS * 16)        // bytes of one time buffer of tuples (8 / 16 KB): the buffers are aligned to it
#define R2_PARW (3 + R2_WIDE)         // 16-byte words per (member, layer) in par_s
#define R2_COLW (3 - R2_WIDE)         // 16-byte words per column in cold_s
#ifndef R2_REG_COMPUTE
#define R2_REG_COMPUTE 240            // setmaxnreg (a) moves registers inside the CTA's LAUNCH allocation (384 x 168) only -- a request
                                      // beyond that pool blocks forever -- and (b) must be the SAME instruction for the four warps of a
                                      // warpgroup (warps 8-11 are the river warps): 256 x 240 + 128 x 24 = 64 512 = 384 x 168
#endif
#ifndef R2_REG_RIVER
#define R2_REG_RIVER 24
#endif
#ifdef DCP_RES2_DEBUG                  // per-warp busy-cycle counters (tuning builds only: two registers and a branch per step)
#define R2_DBG (P.dbg != nullptr)
#else
#define R2_DBG false
#endif
#ifndef R2_LONG
#define R2_LONG 2                     // x R2_NACC entries per round in a thread's first (longest) column; full-length bench on
#endif                                // one box (gpurun_out/ab2.log): 1 / 2 / 3 = 8.264 / 8.338 / 8.238e10 units/s
#ifndef R2_NACC
#define R2_NACC 4                     // partial sums per member slot in the gather
#endif
#ifndef R2_GATHER_UNROLL
#define R2_GATHER_UNROLL 1            // x R2_NACC entries per iteration of the gather loop.  Measured on C2 (gpurun_out/y1,y2_variants.log):
                                      // 1/2/4/8/16 entries per iteration = 4.86/5.38/5.52/5.48/5.27e10 units/s: 4 balance the
                                      // loads in flight against the size of the loop body (instruction-cache misses)
#endif
#if DCP_RES2_GB != 1
#error "the gather walks its list one entry at a time (DCP_RES2_GB == 1)"
#endif
#define R2_S 2                        // member slots per team
#define R2_H 4                        // columns per compute thread
#ifndef R2_BC
#define R2_BC 2                       // columns per physics block: R2_BC x 2 member slots = 4 interleaved dependency chains
                                      // (4 columns = 8 chains spill and lose 17 %: 6.56e10 vs 7.93e10, profiles/r2_ab_routing_designs.txt)
#endif
#ifndef R2_TAB
#define R2_TAB (R2_WIDE ? 2 : 1)      // table-driven exp / log in the lean step (dcp_fastmath.cuh): 1 = tables in shared memory (2.5 KB),
#endif                                // 2 = read through L1 from global memory (the wide build has no shared memory left), 0 = polynomials
#ifndef R2_CARRY_E1
#define R2_CARRY_E1 1                 // carry exp(sd/m2) from step to step instead of recomputing it (dcp_physics.cuh)
#endif

struct R2Smem {                       // team-private views unless noted
    double2 *qbf_s;       // [2][512]        baseflow tuples (2 member slots), double buffered in time
    double2 *con_s;       // [2][512]        river contributions by slot, double buffered in time
    double2 *cst_s;       // [2][4][128]     {q0, q2} per (member slot, column slot j, team thread); m2 = szm/cos(mslp) and its
                          //                 reciprocal are rebuilt from the member's {szm, 1/szm} and the column's {cos, 1/cos}:
                          //                 two DMULs instead of an LDS.128 per unit and step, and 32 KB less shared memory
    double2 *par_s;       // [Mv][npt][3]    {srmax, 1/srmax}, {td/dt, smax}, {szm, dtt/szm}
    double *forc_s;       // [2][Tt*(ngp+nge)]
    uint64_t *bars;       // [2] forcing tiles (TMA)
    int4 *job_s;          // [Mv * jobs_per_member]  routing jobs of the group: {ring offset, delay, taps, histogram offset}
    double *acc_s;        // [Mv * nriv][6]          objective sums sse, lsse, sq, sqq, sqo and the non-finite flag
    // shared by both teams (read-only after setup)
    double *ellw_s;       // [ell_len]       packed gather list (weight | source tuple index)
    int4 *coli_s;         // [4][128]        per (column slot j, team thread): packed indices
    double2 *cold_s;      // [3][4][128]     word 0: {ac, 1/ac}, 1: {cq, cos(mslp)}, 2: {1/cos(mslp), -}; thread-fastest
                          //                 WIDE: [2][4][256] word 0: {ac, cq}, 1: {cos(mslp), 1/cos(mslp)}
    double *eq0_s;        // WIDE: [4][256]  exp(-st) of the column (q0 = exp(t0dt) * exp(-st), dyna_init_satzone.f90:108)
    double *fmtab_s;      // R2_TAB: [64 + 256] tables of fm_exp_tab / fm_log_tab
    int32_t *rbeg_s;      // [nriv+1]
};

// Source tuple of one gather entry.  The entry's low word carries the tuple's BYTE offset (index << 4) in bits 4..12
// and a time buffer of tuples is 8 KB aligned, so the shared-memory address is ONE LOP3 ((lo & 0x1ff0) | base)
// instead of mask + add + scale.  volatile: ordered against the team barrier (also a volatile asm).
__device__ __forceinline__ double2 lds_tuple(int lo, uint32_t buf_u32) {
    uint32_t a;
#if R2_WIDE        // no alignment slack beside 1024 columns: mask + add (the add folds into the load's register + uniform-register address)
    a = ((uint32_t)lo & (uint32_t)(R2_TUPB - 16)) + buf_u32;
#else
    asm("lop3.b32 %0, %1, %3, %2, 0xEA;" : "=r"(a) : "r"(lo), "r"(buf_u32), "n"(R2_TUPB - 16));
#endif
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}

__device__ __forceinline__ void team_sync(int team) { asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(R2_TT) : "memory"); }

#define gc (StepConst{C.dt, C.dtt, C.inv_dtt, C.ntt, M.fmtab_s, M.fmtab_s + FM_EXP_TAB})

// Completes a unit whose lean step raised `slow` (sd > smax, exponent out of the fast range, non-finite): the generic
// saturated-zone statements (hru_step_lean_fixup -> sat_zone_slow) from the saved inputs.  Kept OUT of the time loop's
// body: it takes the shared-memory addresses of the unit's constant words instead of the constants themselves.
struct R2Fix { double sd, e1, qbf, exac; };
__device__ __forceinline__ double2 lds_d2(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
static __device__ __noinline__ R2Fix r2_fix_unit(uint32_t cst_a, uint32_t par_a, uint32_t cold_a, uint32_t eq0_a, double dtt, double inv_dtt,
                                                 int ntt, double qin, double uz, double exus, double sd0) {
    const double2 p1 = lds_d2(par_a + 16), p2 = lds_d2(par_a + 32);      // {td/dt, smax}, {szm, dtt/szm}
#if R2_WIDE
    const double2 w0 = lds_d2(cold_a), w1 = lds_d2(cold_a + R2_H * R2_TC * 16);          // {ac, cq}, {cos, 1/cos}
    double q2v, e0;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(q2v) : "r"(cst_a));
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(e0) : "r"(eq0_a));
    const double2 c0v = make_double2(lds_d2(par_a + 48).x * e0, q2v);                      // {q0, q2}
    const double2 d0 = make_double2(w0.x, 1.0), d1 = make_double2(w0.y, w1.x);            // 1/ac sits in the gather weights
    const double icos = w1.y;
#else
    (void)eq0_a;
    const double2 c0v = lds_d2(cst_a);                                   // {q0, q2}
    const double2 d0 = lds_d2(cold_a), d1 = lds_d2(cold_a + R2_H * R2_TC * 16);
    const double icos = lds_d2(cold_a + 2 * R2_H * R2_TC * 16).x;
#endif
    HruConst c;
    c.ac = d0.x; c.inv_ac = d0.y;
    c.q0 = c0v.x; c.q2 = c0v.y; c.m2 = p2.x * icos; c.inv_m2 = (d1.y * p2.y) * inv_dtt;   // p2.y = dtt/szm
    c.q1 = c.q0 * d1.y;
    c.smax = p1.y;
    LeanOut o;
    o.uz = uz; o.exus = exus; o.sd0 = sd0;
    double sd = sd0;
    hru_step_lean_fixup(c, StepConst{0.0, dtt, inv_dtt, ntt}, qin, qin, sd, o);
    R2Fix f;
    f.sd = sd; f.qbf = o.qbf; f.exac = o.exac;
    f.e1 = exp(sd * c.inv_m2);                                           // re-derive the carried exponential
    return f;
}

// ---------------------------------------------------------------------------------------------
// Routing + objective terms of a finished TILE of steps [t0, t1) (t1 - t0 <= R2_RT), run by the team's four compute warps
// once per R2_RT steps -- NOT by the river warps every step: measured on C2 (profiles/r2_ab_routing_designs.txt), the
// convolution in the river warps costs 8 % for the per-step FIR, 8 % for the per-step gauge terms and 27 % for both (their
// latency chain then exceeds a compute step), a dedicated routing warp 16 %; here the lanes of a warp are 32 consecutive
// TIMESTEPS of one (member, gauge) pair, every lane busy, no barrier, ~2 % of the compute warps' instructions.
//     job(t)  = sum_k delay_hist_u(k + o) * qout_u(t - d - k)       one job per node u in {i} + upstream tree of gauge i
//     q_i(t)  = (sum of the gauge's jobs) / frac_sub_area             (DTA/dta_route_processing.f90:671-714, dyna_river.f90:57-59)
// The inflows come from the member slot's RING (the river warps store qout(t) there one step behind the compute warps),
// the per-(member, gauge) objective sums live in shared memory: a warp shuffle tree per tile, then one lane adds the tile's
// partial sums (deterministic order: tiles ascending).  Only measures -- and the flows when requested -- leave the chip.
// FAST policy only (no summation-order contract); the EXACT policy routes in dcp_route_fused.cuh.
// ---------------------------------------------------------------------------------------------
#define R2_RT 64                      // == the forcing tile DCP_RES_TT: routing happens between two tiles (ring slack: dcp_api.cu)
struct R2RouteArgs {
    const double *dh, *ring, *qobs, *log_obs, *inv_frac, *obs_eps;
    double *q_user;
    uint32_t *fdc;
    const int32_t *gauge_job_beg;
    uint32_t job_s, acc_s;            // shared-memory addresses
    int Rmask, nriv, njm, n_pairs, m_base, B, nstep, n_t, t_warm;
    double flow_err;
};
static __device__ __noinline__ void r2_route_tile(const R2RouteArgs a, const int t0, const int t1, const int warp, const int lane) {
    for (int pg = warp; pg < a.n_pairs; pg += R2_TC / 32) {                  // (member slot, gauge) pairs over the 4 warps
        const int jm = pg / a.nriv, i = pg - jm * a.nriv;
        const int m = a.m_base + jm;
        if (m >= a.B) continue;                                              // warp-uniform
        const int jb = a.gauge_job_beg[i], je = a.gauge_job_beg[i + 1];
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0;
        bool bad = false;
        for (int tb = t0; tb < t1; tb += 32) {
            const int t = tb + lane;
            if (t < t1) {
                double q_sum = 0.0;
                for (int jj = jb; jj < je; ++jj) {
                    int4 e;
                    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(e.x), "=r"(e.y), "=r"(e.z), "=r"(e.w) : "r"(a.job_s + (jm * a.njm + jj) * 16));
                    const int tq = t - e.y;
                    const int nt = min(e.z, tq + 1);                         // terms before the first timestep do not exist
                    const double *__restrict__ dh = a.dh + e.w;
                    const double *qi = a.ring + e.x;
                    double y = 0.0;
#pragma unroll 4
                    for (int k = 0; k < nt; ++k) y = fma(dh[k], qi[(tq - k) & a.Rmask], y);
                    q_sum += y;
                }
                const double q = q_sum * a.inv_frac[i];
                if (a.q_user) a.q_user[((size_t)m * a.nstep + t) * a.nriv + i] = q;
                if (a.fdc && t >= a.t_warm) atomicAdd(a.fdc + ((size_t)m * a.nriv + i) * DCP_FDC_BINS + dcp_fdc_bin(q), 1u);
                const bool fin = isfinite(q);
                bad |= !fin;
                if (a.qobs != nullptr && t >= a.t_warm && t < a.n_t) {
                    const double o = a.qobs[(size_t)t * a.nriv + i];
                    if (o != a.flow_err && fin) {
                        const double eps = a.obs_eps[i];
                        const double d = q - o;
                        const double dl = LEAN_LOG(q > eps ? q : eps) - a.log_obs[(size_t)t * a.nriv + i];   // ln(max(obs, eps)): member independent
                        s0 = fma(d, d, s0); s1 = fma(dl, dl, s1); s2 += q; s3 = fma(q, q, s3); s4 = fma(q, o, s4);
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o); s3 += __shfl_xor_sync(0xffffffffu, s3, o);
            s4 += __shfl_xor_sync(0xffffffffu, s4, o);
        }
        const bool any_bad = __any_sync(0xffffffffu, bad);
        if (lane == 0) {
            const uint32_t ad = a.acc_s + pg * 48;
            double c0, c1, c2, c3, c4;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(c0), "=d"(c1) : "r"(ad));
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(c2), "=d"(c3) : "r"(ad + 16));
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(c4) : "r"(ad + 32));
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(ad), "d"(c0 + s0), "d"(c1 + s1) : "memory");
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(ad + 16), "d"(c2 + s2), "d"(c3 + s3) : "memory");
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(ad + 32), "d"(c4 + s4) : "memory");
            if (any_bad) asm volatile("st.shared.f64 [%0], %1;" ::"r"(ad + 40), "d"(1.0) : "memory");
        }
    }
}

// ---------------------------------------------------------------------------------------------
// compute role: 4 warps per team, thread = 4 HRU columns x 2 member slots
// ---------------------------------------------------------------------------------------------
// ONEPAR: one parameter layer and one member sub-group per team (C2): the member-parameter words sit at fixed
// addresses, so their loads are shared by the columns of a block
template <bool AET, bool ONEPAR>
__device__ __forceinline__ void r2_compute(const DevCatchment &C, const DevBatch &W, const Res2Plan &P, const R2Smem &M,
                                           const int team, const int tt_id) {
    constexpr int H = R2_H, S = R2_S;
    const int B = W.B, npt = C.npt;
    const int forc_stride = P.Tt * (C.ngp + C.nge);
    const int par_kstride = npt * R2_PARW;
    const int ntiles = (C.nstep + P.Tt - 1) / P.Tt;
    const int Mv = P.G * S;                                  // members per team group
    uint32_t tiles_done = 0;
    long long busy = 0;

    for (int vg = blockIdx.x * R2_TEAMS + team; vg < P.n_groups; vg += gridDim.x * R2_TEAMS) {
        const int m_base = vg * Mv;
        team_sync(team);                                     // previous group fully done with shared memory
        auto route_args = [&]() {
            R2RouteArgs a;
            a.dh = W.dh; a.ring = W.qout; a.qobs = C.qobs; a.log_obs = P.log_obs; a.inv_frac = P.inv_frac; a.obs_eps = P.obs_eps;
            a.q_user = W.q_user; a.fdc = W.fdc; a.gauge_job_beg = P.gauge_job_beg;
            a.job_s = smem_u32(M.job_s); a.acc_s = smem_u32(M.acc_s);
            a.Rmask = W.Rmask; a.nriv = C.nriv; a.njm = P.jobs_per_member; a.n_pairs = Mv * C.nriv; a.m_base = m_base; a.B = B;
            a.nstep = C.nstep; a.n_t = C.nstep < C.nstep_obs ? C.nstep : C.nstep_obs; a.t_warm = C.t_warm; a.flow_err = C.flow_err;
            return a;
        };

        // ---- load the group's initial state (written by k_init_members) ----
        double sd[H][S], suz[H][S], srz[H][S], aet[H][S];
#if R2_CARRY_E1
        double e1s[H][S];
#endif
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const int4 ci = M.coli_s[j * R2_TC + tt_id];
            const int tup = ci.z & 0xffff, g = (ci.y >> 16) & 0x7fff;
            const int ia = P.col_ia[j * R2_TC + tt_id];
            const double2 cw1 = M.cold_s[R2_H * R2_TC + j * R2_TC + tt_id];
            const double cosm = R2_WIDE ? cw1.x : cw1.y;
            const int ipar = ia >= 0 ? C.hru[ia].ipar : 0;
#pragma unroll
            for (int k = 0; k < S; ++k) {
                const int m = m_base + g * S + k;
                sd[j][k] = 1.0; suz[j][k] = 0.0; srz[j][k] = 0.0; aet[j][k] = 0.0;
                double qb0 = 0.0, c_q0 = 1.0, c_q2 = 1.0, c_im2 = 1.0;
                if (ia >= 0 && m < B) {
                    const size_t o = (size_t)ia * B + m;
                    sd[j][k] = W.sd[o]; suz[j][k] = W.suz[o]; srz[j][k] = W.srz[o];
                    c_q0 = W.q0[o]; c_q2 = W.q2[o];
                    const double szm = W.par[(0 * npt + ipar) * (size_t)B + m];
                    c_im2 = cosm * (1.0 / szm);              // 1/m2 exactly as the time loop rebuilds it (par_s x cold_s)
                    qb0 = W.qbf0[o];
                }
#if R2_CARRY_E1
                e1s[j][k] = exp(sd[j][k] * c_im2);           // carried exp(sd/m2), see hru_step_lean<true>
#endif
                reinterpret_cast<double *>(M.qbf_s + tup)[k] = qb0;                 // time buffer 0 = step 0's "old" qbf
                reinterpret_cast<double *>(M.qbf_s + R2_COLS + tup)[k] = 0.0;
#if R2_WIDE
                reinterpret_cast<double *>(M.cst_s)[(k * H + j) * R2_TC + tt_id] = c_q2;   // q0 is rebuilt from exp(t0dt) and exp(-st)
                (void)c_q0;
#else
                M.cst_s[(k * H + j) * R2_TC + tt_id] = make_double2(c_q0, c_q2);    // own-column table: [slot][column j][thread]
#endif
            }
        }
        for (int i = tt_id; i < Mv * npt; i += R2_TC) {      // member parameters -> shared
            const int jm = i / npt, l = i % npt;
            const int m = m_base + jm;
            double srmax = 1.0, td = 1.0, smax = 1.0, szm = 1.0;
            if (m < B) {
                szm = W.par[(0 * npt + l) * (size_t)B + m];
                srmax = W.par[(2 * npt + l) * (size_t)B + m];
                td = W.par[(5 * npt + l) * (size_t)B + m];
                smax = W.par[(6 * npt + l) * (size_t)B + m];
            }
            M.par_s[i * R2_PARW + 0] = make_double2(srmax, 1.0 / srmax);
            M.par_s[i * R2_PARW + 1] = make_double2(td / C.dt, smax);     // folded constants of hru_step_lean<., true>
            M.par_s[i * R2_PARW + 2] = make_double2(szm, C.dtt / szm);
#if R2_WIDE
            M.par_s[i * R2_PARW + 3] = make_double2(m < B ? exp(W.t0dt[l * (size_t)B + m]) : 1.0, 0.0);   // exp(t0dt): q0 = exp(t0dt) * exp(-st)
#endif
        }
        team_sync(team);

        // first batch of the thread's gather list (time-invariant; re-issued before every barrier so that its
        // latency is hidden behind the wait)
        const double *ent0 = M.ellw_s + (M.coli_s[tt_id].x & 0xfffff);
        constexpr int GB = DCP_RES2_GB;                      // gather batch (entries in flight per tuple round)
        double wv[GB];
#pragma unroll
        for (int u = 0; u < GB; ++u) wv[u] = ent0[u * 32];

        int t = 0, routed = 0;
        // one pass per forcing tile + a last pass without steps: every pass ends by routing the steps whose river inflows are
        // complete (the river warps run one step behind), OUTSIDE the step loop -- a call inside it costs 6 % (register
        // allocation of the hot basic blocks), between two tiles nothing measurable
#pragma unroll 1
        for (int tile = 0; tile <= ntiles; ++tile) {
            int r_hi = C.nstep;
            if (tile == ntiles) team_sync(team);             // the river warps have stored qout(nstep - 1)
            else {
            const uint32_t b = tiles_done & 1;
            mbar_wait(&M.bars[b], (tiles_done >> 1) & 1);    // forcing tile landed (TMA, issued by the team's river warp)
            const double *forc_t = M.forc_s + b * forc_stride;
            const int t_end = min(C.nstep, t + P.Tt);
            for (int tt = 0; t < t_end; ++t, ++tt) {
                const long long c0 = R2_DBG ? clock64() : 0;
                const double2 *qbf_cur = M.qbf_s + (t & 1) * R2_COLS;
                const uint32_t qcur_u32 = smem_u32(qbf_cur);         // 8 KB aligned (kernel entry), see lds_tuple
                double2 *qbf_nxt = M.qbf_s + ((t & 1) ^ 1) * R2_COLS;
                double2 *con = M.con_s + (t & 1) * R2_COLS;
                // ---- dynamic_dist gather (dyna_dynamic_dist.f90:49-52): the thread's four columns form one flat list
                // (two columns are gathered ahead of each physics block); the next batch's entries are in flight while
                // this batch's tuples are fetched ----
                const double *ep = ent0;
#pragma unroll
                for (int jp = 0; jp < H; jp += R2_BC) {      // R2_BC columns = 2 R2_BC (HRU, member) units per basic block
                    int4 ci[H];
                    double qin[H][S];
#pragma unroll
                    for (int j = jp; j < jp + R2_BC; ++j) {
                        ci[j] = M.coli_s[j * R2_TC + tt_id];
                        const int nb = (ci[j].x >> 20) / GB;                            // K / batch, warp-uniform
                        // R2_NACC independent partial sums per member slot: a row of K entries is K dependent DFMAs per
                        // slot otherwise (8.9 cycles each), the longest chain of the gather phase
                        double qa[R2_NACC][2];
#pragma unroll
                        for (int u = 0; u < R2_NACC; ++u) { qa[u][0] = 0.0; qa[u][1] = 0.0; }
                        constexpr int GU = R2_GATHER_UNROLL;
                        int bb = 0;
#if R2_LONG > 1                       // wide rounds for the thread's first (longest) column only: a round costs two shared-memory
                                      // round trips whatever its width (wide rounds for EVERY column grow the loop body and lose 9 %)
                        if (j == 0) {
#pragma unroll 1
                            for (; bb + R2_LONG * R2_NACC <= nb; bb += R2_LONG * R2_NACC) {
#pragma unroll
                                for (int u = 0; u < R2_LONG * R2_NACC; ++u) {
                                    const double wc = wv[0];
                                    const double2 qv = lds_tuple(__double2loint(wc), qcur_u32);
                                    ep += 32;
                                    wv[0] = ep[0];
                                    qa[u % R2_NACC][0] = fma(qv.x, wc, qa[u % R2_NACC][0]);
                                    qa[u % R2_NACC][1] = fma(qv.y, wc, qa[u % R2_NACC][1]);
                                }
                            }
                        }
#endif
#pragma unroll GU
                        for (; bb + R2_NACC <= nb; bb += R2_NACC) {
#pragma unroll
                            for (int u = 0; u < R2_NACC; ++u) {
                                const double wc = wv[0];
                                const double2 qv = lds_tuple(__double2loint(wc), qcur_u32);
                                ep += 32;
                                wv[0] = ep[0];
                                qa[u][0] = fma(qv.x, wc, qa[u][0]); qa[u][1] = fma(qv.y, wc, qa[u][1]);
                            }
                        }
#pragma unroll 1
                        for (; bb < nb; ++bb) {
                            const double wc = wv[0];
                            const double2 qv = lds_tuple(__double2loint(wc), qcur_u32);
                            ep += 32;
                            wv[0] = ep[0];
                            qa[0][0] = fma(qv.x, wc, qa[0][0]); qa[0][1] = fma(qv.y, wc, qa[0][1]);
                        }
                        double q0 = qa[0][0], q1 = qa[0][1];
#pragma unroll
                        for (int u = 1; u < R2_NACC; ++u) { q0 += qa[u][0]; q1 += qa[u][1]; }
                        qin[j][0] = q0; qin[j][1] = q1;
                    }
                    // ---- root zone, unsaturated zone, saturated zone: 4 interleaved dependency chains ----
                    LeanOut lo[R2_BC][S];
                    double cq[R2_BC];
                    bool any_slow = false;
#pragma unroll
                    for (int jj = 0; jj < R2_BC; ++jj) {
                        const int j = jp + jj;
#if R2_WIDE
                        const double2 w0 = M.cold_s[j * R2_TC + tt_id];      // {ac, cq}
                        const double2 w1 = M.cold_s[R2_H * R2_TC + j * R2_TC + tt_id];      // {cos, 1/cos}
                        const double eq0 = M.eq0_s[j * R2_TC + tt_id];
                        const double2 d0 = make_double2(w0.x, 1.0);          // 1/ac is folded into the gather weights
                        const double2 d1 = make_double2(w0.y, w1.x);
                        const double icos = w1.y;
#else
                        const double2 d0 = M.cold_s[j * R2_TC + tt_id];      // {ac, 1/ac}
                        const double2 d1 = M.cold_s[R2_H * R2_TC + j * R2_TC + tt_id];      // {cq, cos}
                        const double icos = M.cold_s[2 * R2_H * R2_TC + j * R2_TC + tt_id].x;
#endif
                        cq[jj] = d1.x;
                        const double p = forc_t[tt * C.ngp + (ci[j].y & 0xff)];
                        const double epp = forc_t[P.Tt * C.ngp + tt * C.nge + ((ci[j].y >> 8) & 0xff)];   // already max(ep, 0)
#pragma unroll
                        for (int k = 0; k < S; ++k) {
#if R2_WIDE
                            const double2 c0v = make_double2(M.par_s[(ONEPAR ? 0 : (ci[j].w & 0xffff)) + k * par_kstride + 3].x * eq0,
                                                             reinterpret_cast<const double *>(M.cst_s)[(k * H + j) * R2_TC + tt_id]);
#else
                            const double2 c0v = M.cst_s[(k * H + j) * R2_TC + tt_id];
#endif
                            const double2 p0 = M.par_s[(ONEPAR ? 0 : (ci[j].w & 0xffff)) + k * par_kstride];
                            const double2 p1 = M.par_s[(ONEPAR ? 0 : (ci[j].w & 0xffff)) + k * par_kstride + 1];
                            const double2 p2 = M.par_s[(ONEPAR ? 0 : (ci[j].w & 0xffff)) + k * par_kstride + 2];
                            HruConst c;
                            c.ac = d0.x; c.inv_ac = d0.y;
                            c.q0 = c0v.x; c.q2 = c0v.y; c.m2 = p2.x * icos; c.dtt_inv_m2 = d1.y * p2.y;   // dyna_satzone_analyticalsolver.f90:52
                            c.q1 = c.q0 * d1.y;                                         // dyna_satzone_analyticalsolver.f90:49
                            c.srmax = p0.x; c.inv_srmax = p0.y; c.td_dt = p1.x; c.smax = p1.y;
                            c.ims_rz = 1; c.ims_uz = 1;
                            HruState st;
                            st.sd = sd[j][k]; st.suz = suz[j][k]; st.srz = srz[j][k]; st.qint1 = 0.0;
#if R2_CARRY_E1
                            st.e1 = e1s[j][k];
#endif
                            hru_step_lean<R2_CARRY_E1 != 0, true, R2_TAB != 0>(c, gc, p, epp, qin[j][k], st, lo[jj][k]);
                            sd[j][k] = st.sd; suz[j][k] = st.suz; srz[j][k] = st.srz;  // ntt == 1: qint1 is not carried
#if R2_CARRY_E1
                            e1s[j][k] = st.e1;
#endif
                            any_slow |= lo[jj][k].slow;
                        }
                    }
                    if (__builtin_expect(any_slow, 0)) {                                                     // rare: sd > smax, out-of-range exponent
#pragma unroll
                        for (int jj = 0; jj < R2_BC; ++jj)
#pragma unroll
                            for (int k = 0; k < S; ++k)
                                if (lo[jj][k].slow) {
                                    const int j = jp + jj;
                                    // out of line (the loop body must stay small, see R2_GATHER_UNROLL): the unit's constants are
                                    // re-read from shared memory inside
                                    const R2Fix f = r2_fix_unit(R2_WIDE ? smem_u32(reinterpret_cast<double *>(M.cst_s) + (k * H + j) * R2_TC + tt_id)
                                                                        : smem_u32(M.cst_s + (k * H + j) * R2_TC + tt_id),
                                                                smem_u32(M.par_s + (ci[j].w & 0xffff) + k * par_kstride),
                                                                smem_u32(M.cold_s + j * R2_TC + tt_id), smem_u32(M.eq0_s + j * R2_TC + tt_id), C.dtt, C.inv_dtt, C.ntt,
                                                                qin[j][k], lo[jj][k].uz, lo[jj][k].exus, lo[jj][k].sd0);
                                    sd[j][k] = f.sd; lo[jj][k].qbf = f.qbf; lo[jj][k].exac = f.exac;
#if R2_CARRY_E1
                                    e1s[j][k] = f.e1;
#endif
                                }
                    }
                    // publish: both loads of the old baseflow first (straight-line, so that they overlap), then the stores.
                    // Idle columns (bit 31 of .y) take part except for the tuple store: their tuples must stay 0.0 (they are
                    // the zero-weight padding targets of the gather list) and their contribution slots lie past every river.
                    double2 qold[R2_BC];
#pragma unroll
                    for (int jj = 0; jj < R2_BC; ++jj) qold[jj] = qbf_cur[ci[jp + jj].z & 0xffff];  // baseflow at the start of the step
#pragma unroll
                    for (int jj = 0; jj < R2_BC; ++jj) {
                        const int j = jp + jj;
                        const int tup = ci[j].z & 0xffff, ctup = (int)((uint32_t)ci[j].w >> 16);
                        if (ci[j].y >= 0) qbf_nxt[tup] = make_double2(lo[jj][0].qbf, lo[jj][1].qbf);
                        con[ctup] = make_double2(fma(cq[jj], qold[jj].x, lo[jj][0].exac), fma(cq[jj], qold[jj].y, lo[jj][1].exac));
                        if (AET) {
#pragma unroll
                            for (int k = 0; k < S; ++k) aet[j][k] = aet[j][k] + lo[jj][k].ea;
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < GB; ++u) wv[u] = ent0[u * 32];                     // next step's first gather batch
                if (R2_DBG) busy += clock64() - c0;
                team_sync(team);
            }
            ++tiles_done;
            r_hi = t - 1;                                    // qout of every step < t - 1 ... t - 2 is in the rings; step t - 1 follows next pass
            }
            if (r_hi > routed) r2_route_tile(route_args(), routed, r_hi, tt_id >> 5, tt_id & 31);
            routed = r_hi > routed ? r_hi : routed;
        }
        __syncwarp();
        for (int pg = (tt_id >> 5); pg < Mv * C.nriv; pg += R2_TC / 32) {               // the pair's warp wrote its sums: lane 0 hands them over
            const int jm = pg / C.nriv, i = pg - jm * C.nriv;
            const int m = m_base + jm;
            if ((tt_id & 31) == 0 && m < B) {
                const double *acc = M.acc_s + pg * 6;
                const size_t a = i * (size_t)B + m;
                W.sse[a] = acc[0]; W.lsse[a] = acc[1]; W.sq[a] = acc[2]; W.sqq[a] = acc[3]; W.sqo[a] = acc[4];
                if (acc[5] != 0.0) W.bad[a] = 1;
            }
        }
        if (R2_DBG && (tt_id & 31) == 0) P.dbg[blockIdx.x * 12 + team * 4 + (tt_id >> 5)] = busy;
        if (AET) {
#pragma unroll
            for (int j = 0; j < H; ++j) {
                const int ia = P.col_ia[j * R2_TC + tt_id];
                const int g = (M.coli_s[j * R2_TC + tt_id].y >> 16) & 0x7fff;
#pragma unroll
                for (int k = 0; k < S; ++k) {
                    const int m = m_base + g * S + k;
                    if (ia >= 0 && m < B) W.aet[(size_t)ia * B + m] = aet[j][k];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// river role: 2 warps per team, one step behind the compute warps.  A group of T lanes adds up one (member sub-group,
// river) range of contribution tuples: qout_r(tp) (dyna_river.f90:43), stored into the member slot's inflow RING -- the last
// `hist_len` (+ one routing tile) values per river are all the routing ever needs, so no series of length nstep exists
// unless the caller asks for the flows.  At the start of a group the river lanes also lay out the group's routing jobs
// (the histogram geometry depends on each member's channel velocity) and clear its objective sums; lane 0 drives the
// team's forcing TMA.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void r2_river(const DevCatchment &C, const DevBatch &W, const Res2Plan &P, const R2Smem &M,
                                         const int team, const int rl) {
    constexpr int S = R2_S;
    const int nac_pad = P.nac_pad, B = W.B, nriv = C.nriv;
    const int forc_stride = P.Tt * (C.ngp + C.nge);
    const int ntiles = (C.nstep + P.Tt - 1) / P.Tt;
    const uint32_t rain_bytes = (uint32_t)(P.Tt * C.ngp * sizeof(double));
    const uint32_t pet_bytes = (uint32_t)(P.Tt * C.nge * sizeof(double));
    const int Mv = P.G * S;
    uint32_t tiles_done = 0;
    const int T = P.team, sub = rl & (T - 1);
    const int sums_per_pass = R2_TR / T;
    const int sl_base = (blockIdx.x * R2_TEAMS + team) * Mv; // ring slots of this team (one per member in flight)
    const int njm = P.jobs_per_member, n_jobs = Mv * njm;
    long long busy = 0;

    for (int vg = blockIdx.x * R2_TEAMS + team; vg < P.n_groups; vg += gridDim.x * R2_TEAMS) {
        const int m_base = vg * Mv;
        team_sync(team);
        if (rl == 0) {                                       // forcing tile 0 of this group
            const uint32_t b = tiles_done & 1;
            fence_proxy_async();
            mbar_expect_tx(&M.bars[b], rain_bytes + pet_bytes);
            tma_load_1d(M.forc_s + b * forc_stride, C.rain, rain_bytes, &M.bars[b]);
            tma_load_1d(M.forc_s + b * forc_stride + P.Tt * C.ngp, P.pet_pos, pet_bytes, &M.bars[b]);
        }
        // routing jobs of this group's members: the histogram geometry depends on the member's channel velocity
        for (int j = rl; j < n_jobs; j += R2_TR) {
            const int jm = j / njm, jj = j - jm * njm;
            const int m = m_base + jm;
            int4 e = make_int4(0, 0, 0, 0);
            if (m < B) {
                const int u = P.job_node[jj], i = P.job_gauge[jj];
                const int riv = C.node_input[u];
                const int bins = W.hbins[u * (size_t)B + m];
                const int d = W.hstart[u * (size_t)B + m] - W.hstart[i * (size_t)B + m];    // hist start of u past the gauge's sum_index
                const int o = d < 0 ? -d : 0;                // column inside u's histogram: partial sum from tap o on
                if (riv >= 0 && bins - o > 0) {
                    e.x = ((sl_base + jm) * nriv + riv) * W.R;
                    e.y = d < 0 ? 0 : d;
                    e.z = bins - o;
                    e.w = (int)(((size_t)m * C.nnode + u) * W.Lcap) + o;
                }
            }
            M.job_s[j] = e;
        }
        for (int a = rl; a < Mv * nriv * 6; a += R2_TR) M.acc_s[a] = 0.0;
        team_sync(team);

        // river sums, routing and objective terms of step tp
        auto flush_step = [&](int tp) {
            const double2 *con = M.con_s + (tp & 1) * R2_COLS;
#pragma unroll 1
            for (int pass = 0; pass < P.n_pass; ++pass) {
                const int sum_id = pass * sums_per_pass + rl / T;
                const bool live = sum_id < P.n_teams;
                const int g = live ? sum_id / nriv : 0, r = live ? sum_id - g * nriv : 0;
                const int beg = M.rbeg_s[r], end = live ? M.rbeg_s[r + 1] : beg;
                const double2 *cg = con + g * nac_pad;
                double a0 = 0.0, a1 = 0.0;
#pragma unroll 4
                for (int p = beg + sub; p < end; p += T) { const double2 x = cg[p]; a0 += x.x; a1 += x.y; }
                for (int o = T >> 1; o > 0; o >>= 1) {
                    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
                    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
                }
                if (live && sub == 0) {
                    const size_t tq = (size_t)(tp & W.Rmask);
                    W.qout[((size_t)(sl_base + g * S + 0) * nriv + r) * W.R + tq] = a0;
                    W.qout[((size_t)(sl_base + g * S + 1) * nriv + r) * W.R + tq] = a1;
                }
            }
        };

        // one loop, one copy of flush_step in the instruction stream (the river code shares the SM's instruction cache with
        // the compute warps' 30 KB loop body): iteration t sums / routes step t-1 while the compute warps run step t
#pragma unroll 1
        for (int t = 0; t <= C.nstep; ++t) {
            if (t < C.nstep && (t & (P.Tt - 1)) == 0) {      // first step of a forcing tile (Tt is a power of two)
                if (t > 0) ++tiles_done;
                if (rl == 0 && t + P.Tt < C.nstep) {
                    // prefetch the next tile into the other buffer (the team left it at the last barrier)
                    const uint32_t nb = (tiles_done & 1) ^ 1;
                    const size_t t_next = (size_t)t + P.Tt;
                    fence_proxy_async();
                    mbar_expect_tx(&M.bars[nb], rain_bytes + pet_bytes);
                    tma_load_1d(M.forc_s + nb * forc_stride, C.rain + t_next * C.ngp, rain_bytes, &M.bars[nb]);
                    tma_load_1d(M.forc_s + nb * forc_stride + P.Tt * C.ngp, P.pet_pos + t_next * C.nge, pet_bytes, &M.bars[nb]);
                }
            }
            const long long c0 = R2_DBG ? clock64() : 0;
            if (t > 0) flush_step(t - 1);
            if (R2_DBG) busy += clock64() - c0;
            if (t < C.nstep) team_sync(team);
        }
        ++tiles_done;
        team_sync(team);                                     // qout of the last step is in the rings: the compute warps route the tail
        if (R2_DBG && (rl & 31) == 0) P.dbg[blockIdx.x * 12 + 8 + team * 2 + (rl >> 5)] = busy;
    }
}

template <bool AET, bool ONEPAR>
__global__ void __launch_bounds__(R2_THREADS, 1) R2_KERNEL(DevCatchment C, DevBatch W, Res2Plan P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    // warps 0-3: team 0 compute, 4-7: team 1 compute, 8-9: team 0 river, 10-11: team 1 river
    const bool is_compute = tid < R2_CT;
    const int team = is_compute ? tid / R2_TC : (tid - R2_CT) / R2_TR;
    R2Smem M;
    // team blocks start on an 8 KB boundary of the shared window (their first array is the tuple buffer pair, whose
    // 8 KB halves are addressed by OR-ing the tuple offset into the base: lds_tuple); the plan reserves the slack
    const uint32_t pad = R2_WIDE ? 0u : ((uint32_t)R2_TUPB - ((smem_u32(smem) + (uint32_t)P.off_team) & (R2_TUPB - 1u))) & (R2_TUPB - 1u);
    unsigned char *tb = smem + P.off_team + pad + (size_t)team * P.team_bytes;
    M.qbf_s = (double2 *)(tb + P.off_qbf);
    M.con_s = (double2 *)(tb + P.off_con);
    M.cst_s = (double2 *)(tb + P.off_cst);
    M.par_s = (double2 *)(tb + P.off_par);
    M.forc_s = (double *)(tb + P.off_forc);
    M.bars = (uint64_t *)(tb + P.off_bar);
    M.job_s = (int4 *)(tb + P.off_job);
    M.acc_s = (double *)(tb + P.off_acc);
    M.ellw_s = (double *)(smem + P.off_ellw);
    M.coli_s = (int4 *)(smem + P.off_coli);
    M.cold_s = (double2 *)(smem + P.off_cold);
    M.eq0_s = (double *)(smem + P.off_eq0);
    M.fmtab_s = R2_TAB == 2 ? const_cast<double *>(P.fm_tab) : (double *)(smem + P.off_fmtab);
    if (R2_TAB == 1) for (int i = tid; i < FM_EXP_TAB + 2 * FM_LOG_TAB; i += R2_THREADS) M.fmtab_s[i] = P.fm_tab[i];
    M.rbeg_s = (int32_t *)(smem + P.off_rbeg);

    for (int e = tid; e < P.ell_len; e += R2_THREADS) M.ellw_s[e] = P.ell_w[e];
    for (int i = tid; i < R2_H * R2_TC; i += R2_THREADS) {
        M.coli_s[i] = P.col_i[i];
    }
    for (int i = tid; i < R2_COLW * R2_H * R2_TC; i += R2_THREADS) M.cold_s[i] = P.col_d[i];
    if (R2_WIDE) for (int i = tid; i < R2_H * R2_TC; i += R2_THREADS) M.eq0_s[i] = P.eq0[i];
    for (int i = tid; i <= C.nriv; i += R2_THREADS) M.rbeg_s[i] = P.rbeg[i];
    if (!is_compute && (tid - R2_CT) % R2_TR == 0) {
        mbar_init(&M.bars[0], 1);
        mbar_init(&M.bars[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    static_assert(R2_CT * R2_REG_COMPUTE + R2_TEAMS * R2_TR * R2_REG_RIVER <= R2_THREADS * 168,
                  "register split exceeds the CTA's launch allocation: setmaxnreg.inc would never return");
    if (is_compute) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R2_REG_COMPUTE));
        if (P.phase_delay < 0 && team == 1) return;          // tuning experiment: one team only (half the members are skipped)
        if (team == 1 && P.phase_delay > 0) {                // start the second team out of phase with the first
            const long long t0 = clock64();
            while (clock64() - t0 < P.phase_delay) {}
        }
        r2_compute<AET, ONEPAR>(C, W, P, M, team, tid % R2_TC);
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R2_REG_RIVER));
        if (P.phase_delay < 0 && team == 1) return;
        r2_river(C, W, P, M, team, (tid - R2_CT) % R2_TR);
    }
}
#undef gc

cudaError_t R2_LAUNCH(const DevCatchment &C, const DevBatch &W, const Res2Plan &P, bool aet,
                                           int grid, cudaStream_t s) {
    const bool onepar = C.npt == 1 && P.G == 1;
    void (*k)(DevCatchment, DevBatch, Res2Plan) =
        aet ? (onepar ? R2_KERNEL<true, true> : R2_KERNEL<true, false>)
            : (onepar ? R2_KERNEL<false, true> : R2_KERNEL<false, false>);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, P.smem_bytes);
    if (e != cudaSuccess) return e;
    k<<<grid, R2_THREADS, P.smem_bytes, s>>>(C, W, P);
    return cudaGetLastError();
}
