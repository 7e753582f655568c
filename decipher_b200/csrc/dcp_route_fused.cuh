// dcp_route_fused.cuh — river routing + objective sums INSIDE the persistent kernels.
//
// The reference routes every timestep as it is produced: river() adds histogram x inflow to a flow matrix, reads the
// gauge columns and shifts the matrix one step (RRMODEL/dyna_river.f90:42-59, DTA/dta_route_processing.f90:671-727).
// Written as the equivalent FIR that needs, per simulated member, only the last `hist_len` values of
// the river inflows qout_r(t) and of the routed own-reach series y_n(t): two small RINGS per member in flight, reused
// by the next member the CTA simulates.  Nothing of length nstep is materialised unless the caller asks for the flows.
//
// Roles inside a persistent kernel:
//   river-sum warps ("A")  write qout_r(t) of every finished step into the q ring and hand over every DCP_RT_TILE steps;
//   ONE routing warp ("B") per team / CTA then processes that tile in chunks of 32 steps, lane = timestep:
//       y_n(t)  = sum_k delay_hist_n(k) qout_n(t-k), oldest term first           (route_process_run_step, :671-681)
//       q_i(t)  = (y_i + sum_{u upstream of i} y_u(t - d_iu)) / frac_sub_area    (:685-714, dyna_river.f90:57-59)
//       the objective terms of step t, added in ascending time through warp shuffles (lane s hands step s to the sums)
//   while the river-sum warps already fill the next tile.  Only the measures (and the flows when requested) leave the chip.
//
// Every operation is spelled with round-to-nearest intrinsics (no FMA contraction whatever the TU's -fmad setting), in
// the order of k_route_conv / k_route_gauge / k_objective_sum (dcp_kernels.cu), so under the EXACT policy the fused path
// returns the same bits as the three routing kernels it replaces and as the oracle's shifted flow matrix.  The FAST
// policy fuses the multiply-add of each histogram tap (half the FP64 instructions of the convolution).
#pragma once
#include "dcp_device.cuh"

#define DCP_RT_TILE 64                       // steps between hand-overs A -> B
#define DCP_RT_SLACK (2 * DCP_RT_TILE + 1)   // a ring holds hist_len + this many steps (A fills tile j+1 while B reads tile j)

// y_n(t) for the member in ring slot `sl` (k_route_conv on rings).  FMA = false: multiply and add rounded separately,
// the reference's arithmetic (EXACT policy); FMA = true: one fused operation per tap (FAST policy, 1e-9 contract).
template <bool FMA>
__device__ __forceinline__ double rt_fir(const DevCatchment &C, const DevBatch &W, int m, int sl, int n, int t) {
    const int riv = C.node_input[n];
    const int bins = W.hbins[n * (size_t)W.B + m];
    if (riv < 0 || bins <= 0) return 0.0;
    const double *__restrict__ dh = W.dh + ((size_t)m * C.nnode + n) * W.Lcap;
    const double *qi = W.qout + ((size_t)sl * C.nriv + riv) * W.R;
    int k = bins - 1;
    if (k > t) k = t;                                 // terms before the first timestep do not exist
    double y = 0.0;
#pragma unroll 8
    for (; k >= 0; --k) {
        const double h = dh[k], q = qi[(t - k) & W.Rmask];
        y = FMA ? fma(h, q, y) : __dadd_rn(y, __dmul_rn(h, q));
    }
    return y;
}

// Pulls the taps of node n and the window of its inflow ring that the chunk [t0, t0 + 32) convolves into L1 with one
// request per lane (lane l touches line l of each), so the FIR loop's first touch of every line is not a serial L2 trip.
__device__ __forceinline__ void rt_prefetch(const DevCatchment &C, const DevBatch &W, int m, int sl, int n, int t0) {
    const int lane = threadIdx.x & 31;
    const int riv = C.node_input[n];
    const int bins = W.hbins[n * (size_t)W.B + m];
    if (riv < 0 || bins <= 0) return;
    const double *dh = W.dh + ((size_t)m * C.nnode + n) * W.Lcap;
    const double *qi = W.qout + ((size_t)sl * C.nriv + riv) * W.R;
    for (int e = lane * 16; e < bins; e += 512) asm volatile("prefetch.global.L1 [%0];" ::"l"(dh + e));
    const int first = t0 - (bins - 1);                // oldest step of the window (may be negative: clipped)
    for (int e = lane * 16; e < bins + 31 + 16; e += 512) {
        const int tt = first + e;
        if (tt >= 0 && tt < t0 + 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(qi + (tt & W.Rmask)));
    }
}

// flow_matrix(u, col) at step t (node_at_column of dcp_kernels.cu on rings)
__device__ __forceinline__ double rt_node_at_column(const DevCatchment &C, const DevBatch &W, int m, int sl, int u, int col, int t) {
    const int B = W.B;
    const int hs = W.hstart[u * (size_t)B + m];
    const int bins = W.hbins[u * (size_t)B + m];
    if (bins <= 0) return 0.0;
    const int d = hs - col;
    if (d >= 0) {                                     // column upstream of u's histogram: pure delay
        if (t - d < 0) return 0.0;
        return W.y[((size_t)sl * C.nnode + u) * W.R + ((t - d) & W.Rmask)];
    }
    const int riv = C.node_input[u];                  // column inside (or past) the histogram of u: partial sum, oldest first
    if (riv < 0) return 0.0;
    const int o = -d;
    double y = 0.0;
    int k = bins - 1 - o;
    if (k > t) k = t;
    const double *dh = W.dh + ((size_t)m * C.nnode + u) * W.Lcap;
    const double *qi = W.qout + ((size_t)sl * C.nriv + riv) * W.R;
    for (; k >= 0; --k) y = __dadd_rn(y, __dmul_rn(dh[k + o], qi[(t - k) & W.Rmask]));
    return y;
}

__device__ __forceinline__ double rt_shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// One chunk of up to 32 steps [t0, t0 + nv) of member m (ring slot sl); all 32 lanes of the routing warp call it.
template <bool FMA>
__device__ __forceinline__ void rt_chunk(const DevCatchment &C, const DevBatch &W, int m, int sl, int t0, int nv) {
    const int lane = threadIdx.x & 31;
    const int t = t0 + lane;
    const bool in = lane < nv;
    const int B = W.B;
    for (int n = 0; n < C.nnode; ++n) rt_prefetch(C, W, m, sl, n, t0);
    for (int n = 0; n < C.nnode; ++n) {
        if (in) W.y[((size_t)sl * C.nnode + n) * W.R + (t & W.Rmask)] = rt_fir<FMA>(C, W, m, sl, n, t);
    }
    __syncwarp();                                     // y of this chunk visible to the lanes that read it with a delay
    const int n_t = C.nstep < C.nstep_obs ? C.nstep : C.nstep_obs;
    for (int i = 0; i < C.nriv; ++i) {                // river / output index; flow_output_indexes(i) = i
        double q = 0.0;
        if (in) {
            const int col = W.hstart[i * (size_t)B + m];          // sum_index
            double q_sum = 0.0;
            q_sum = __dadd_rn(q_sum, rt_node_at_column(C, W, m, sl, i, col, t));
            for (int e = C.up_beg[i]; e < C.up_beg[i + 1]; ++e)
                q_sum = __dadd_rn(q_sum, rt_node_at_column(C, W, m, sl, C.up_idx[e], col, t));
            q = __ddiv_rn(q_sum, C.frac_sub_area[i]);
            if (W.q_user) W.q_user[((size_t)m * C.nstep + t) * C.nriv + i] = q;
            if (W.fdc && t >= C.t_warm) atomicAdd(W.fdc + ((size_t)m * C.nriv + i) * DCP_FDC_BINS + dcp_fdc_bin(q), 1u);
        }
        const bool bad = in && !isfinite(q);
        if (__any_sync(0xffffffffu, bad) && lane == 0) W.bad[i * (size_t)B + m] = 1;
        if (C.qobs == nullptr) continue;
        double o = C.flow_err;                        // masked steps carry the missing-data code
        if (in && t >= C.t_warm && t < n_t) o = C.qobs[(size_t)t * C.nriv + i];
        double d2 = 0.0, dl2 = 0.0;
        if (o != C.flow_err) {
            const double eps = C.obs[i].eps;
            const double d = __dsub_rn(q, o);
            d2 = __dmul_rn(d, d);
            const double dl = __dsub_rn(log(q > eps ? q : eps), log(o > eps ? o : eps));
            dl2 = __dmul_rn(dl, dl);
        }
        // ordered sums: lane s hands the terms of step t0 + s to every lane; all lanes carry identical copies
        const size_t a = i * (size_t)B + m;
        double sse = W.sse[a], lsse = W.lsse[a], sq = W.sq[a], sqq = W.sqq[a], sqo = W.sqo[a];
        for (int s = 0; s < nv; ++s) {
            const double os = rt_shfl(o, s);
            sse = __dadd_rn(sse, rt_shfl(d2, s));
            lsse = __dadd_rn(lsse, rt_shfl(dl2, s));
            if (os != C.flow_err) {                   // moment sums over the valid steps only
                const double qs = rt_shfl(q, s);
                sq = __dadd_rn(sq, qs);
                sqq = __dadd_rn(sqq, __dmul_rn(qs, qs));
                sqo = __dadd_rn(sqo, __dmul_rn(qs, os));
            }
        }
        if (lane == 0) { W.sse[a] = sse; W.lsse[a] = lsse; W.sq[a] = sq; W.sqq[a] = sqq; W.sqo[a] = sqo; }
    }
    __syncwarp();
}

// steps [t0, t1) (one hand-over tile) of the members m_base .. m_base + n_mem - 1 held in ring slots sl_base ..
template <bool FMA>
__device__ __forceinline__ void rt_tile(const DevCatchment &C, const DevBatch &W, int m_base, int n_mem, int sl_base, int t0, int t1) {
    for (int jm = 0; jm < n_mem; ++jm) {
        if (m_base + jm >= W.B) break;
        for (int c0 = t0; c0 < t1; c0 += 32) rt_chunk<FMA>(C, W, m_base + jm, sl_base + jm, c0, min(32, t1 - c0));
    }
}
