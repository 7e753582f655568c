// dcp_device.cuh — device-side data model of the ensemble path.
//
// Layout rule: every per-member quantity is stored MEMBER-FASTEST ("ensemble members are mapped
// across threads"): element (x, m) of a [X][B] array lives at x*B + m, so a warp whose lanes are 32
// consecutive members reads/writes one 256-byte line per access, and every per-HRU / per-river /
// per-cell structure (weights, indices, forcing) is warp-uniform and comes through the read-only
// path as a broadcast.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/decipher_b200.h"

// float32 literals of the reference widened to double
#define DCP_F32_1EM10 ((double)0.0000000001f)
#define DCP_F32_1EM6  ((double)0.000001f)
#define DCP_F32_1EM4  ((double)0.0001f)
#define DCP_F32_1EM2  ((double)0.01f)
#define DCP_F32_1EM9  ((double)1e-9f)

// Static per-HRU record (warp-uniform reads). 64 bytes.
struct __align__(16) HruStatic {
    double ac;        // fractional area
    double st;        // topographic index
    double cosm;      // cos(mslp), computed once on the host (dyna_satzone_analyticalsolver.f90:49)
    double w_riv;     // wtrans(ntrans+1)
    int32_t ippt, ipet, ipar, ipriv;      // 0-based here
    int32_t row_beg, row_end;             // destination-major W row (gather form of dynamic_dist)
    int32_t riv_beg, riv_end;             // non-zero entries of rivers(ia, :)
    int32_t ims_rz, ims_uz;
    int32_t pad0, pad1;
};

// One non-zero of the gather matrix: qin(dest) += (qbf(src) * w) * ac(src)   (dyna_dynamic_dist.f90:49-52)
struct __align__(8) WEntry {
    double w;
    double ac_src;
    int32_t src;
    int32_t pad;
};

struct __align__(8) RivEntry {      // rivers(ia, r) != 0
    double share;
    int32_t r;
    int32_t pad;
};

struct ObsStat {                    // per gauge, host-computed once (member independent)
    double mean, sst, eps, lmean, lsst;
    double sum;                     // sum of the valid observations (PBIAS)
    int32_t n;
    int32_t pad;
};

// Everything a kernel needs; passed by value (fits in the 4 KB parameter space).
struct DevCatchment {
    int32_t nac, nriv, npt, nstep, ngp, nge, ntt, nnode, ncell;
    double dt, dtt, log_dt, inv_dtt;
    const HruStatic *hru;           // [nac]
    const WEntry *w;                // gather CSR
    const RivEntry *riv;            // river CSR
    const double *rain, *pet;       // [ngp x nstep], [nge x nstep]
    const double *sum_ac_riv, *qobs_start;   // [nriv]
    const uint8_t *layer_used;      // [npt] some HRU references the layer (srinit clip, dyna_init_satzone.f90:203)
    // routing, member independent parts of route_processing_build_hist
    const int32_t *node_cell_beg, *node_cell_end;  // [nnode] cell range (0-based, end exclusive), beg<0: no cells
    const double *node_reach_dist;  // [nnode]  max(dist) - min(ds_dist)          (dta_route_processing.f90:226)
    const double *node_down_dist;   // [nnode]  min(ds_dist) - global min ds_dist (:238)
    const double *cell_a, *cell_b, *cell_area;     // [ncell]  section_dist, section_dist + cell length, area
    const int32_t *node_input;      // [nnode] river index feeding the node (inverse of node_to_flow_mapping), -1 none
    const int32_t *up_beg;          // [nnode+1] upstream-tree CSR
    const int32_t *up_idx;          // 0-based node indexes
    const double *frac_sub_area;    // [nriv] frac_sub_area(node_to_flow_mapping(iriv))
    // objective
    const double *qobs;             // [nriv x nstep_obs] or null
    const ObsStat *obs;             // [nriv]
    int32_t nstep_obs, t_warm;
    double flow_err;
    // member-independent reach totals of precipitation / PET (dyna_write_output.f90:61-77)
    const double *tot_ppt, *tot_pet, *tot_frac;    // [nriv]
};

// Per-batch workspace (B members, member fastest).
struct DevBatch {
    int32_t B;                      // members in this batch
    int32_t Lcap;                   // bins reserved per node histogram
    int32_t R;                      // length in time of the qout / y buffers
    int32_t Rmask;                  // R-1 when R is a power of two used as a ring; 0x7fffffff when R >= nstep (no wrap)
    double *par;                    // [7][npt][B]  szm,lnto,srmax,srinit,chv,td,smax
    double *t0dt;                   // [npt][B]
    // state [nac][B]
    double *sd, *suz, *srz, *qint1, *qbf0, *qbf1, *q0, *q2, *aet;
    // checked-mode extras [nac][B] (null otherwise)
    double *sum_pptin, *sum_pein, *sumexs, *sumexus, *sdini, *srzini, *qbfini;
    double *riv_sums;               // [3][nriv][B]  sumqb, sumqof, sumqout (checked mode)
    // routing
    double *dh;                     // [B][nnode][Lcap] normalised delay histograms (bins contiguous)
    int32_t *hstart;                // [nnode][B]  node_hist_indexes(:,1)
    int32_t *hbins;                 // [nnode][B]  bin count
    double *qout;                   // [B][nriv][R]   river inflow series, time fastest; fused routing (dcp_route_fused.cuh):
                                    //                [slots][nriv][R] rings, one slot per member in flight
    double *y;                      // [B][nnode][R]  routed own-reach series ([slots][nnode][R] when fused)
    double *q_user;                 // fused routing: the caller's flow output [nriv x nstep x B] (ABI layout) or null
    double *d2, *dl2;               // [B][nriv][R]   per-step objective terms (null without observed flow)
    double *qs;                     // [B][nriv][R]   routed flow series kept for the moment sums (null without observed flow)
    // objective accumulators [nriv][B]
    double *sse, *lsse;
    double *sq, *sqq, *sqo;         // sums over the valid steps of q, q*q, q*qobs (KGE, PBIAS)
    int32_t *bad;                   // [nriv][B] non-finite flow seen
    uint32_t *fdc;                  // [B][nriv][DCP_FDC_BINS] histogram of the routed flow (DCP_OPT_SIGNATURES), else null
    // per member
    int32_t *status;                // [B]
    double *diag;                   // [3][B]
    // stepwise path (dcp_stepwise.cu), null otherwise
    double *qin_ext;                // [nac][B]  redistributed inflow of the current step
    double *qb_term;                // [nq][B]   per river entry: (((dtt*qbf)*w_riv)*rivers)*ac of the current step
    double *qof_term;               // [nac][B]  (exs+exus)*ac of the current step (0 when not positive)
    int32_t *bad_src;               // [B]       member holds a non-finite baseflow (dense redistribute only)
    int32_t *t_dev;                 // [1]       timestep counter read by the graph-replayed launch sets
};

// Flow-duration histogram bin of a routed flow (DCP_OPT_SIGNATURES): cut on the IEEE bits -- exponent and the top three
// mantissa bits -- so that a host restatement bins the same doubles identically.  Bin 0: below 2^-40 (and anything not
// positive, NaN included), last bin: 2^8 and above.
__host__ __device__ inline int dcp_fdc_bin(double q) {
    long long b;
#if defined(__CUDA_ARCH__)
    b = __double_as_longlong(q);
#else
    static_assert(sizeof(long long) == sizeof(double), "");
    __builtin_memcpy(&b, &q, 8);
#endif
    if (!(q > 0.0)) return 0;
    const int key = (int)(b >> 49) - (1023 + DCP_FDC_OCTAVE_LO) * 8;
    return key < 0 ? 0 : (key >= 8 * DCP_FDC_OCTAVES ? DCP_FDC_BINS - 1 : key + 1);
}

// ---- stepwise path: plan of the redistribute step and of the ordered river sums -------------------
struct DensePlan {
    int32_t feasible;
    int32_t dense;                  // 1: W is dense enough for the GEMM form (wd/kt_* valid)
    int32_t Mp, Kp;                 // nac padded to the GEMM tile (128 destinations, 16 sources)
    int32_t nq;                     // non-zero entries of rivers(:, :)
    const double *wd;               // [Mp][Kp] wtrans in destination-major dense form (0 where no link)
    const double *acs;              // [Kp]     ac(src)
    const int32_t *kt_beg;          // [Mp/128 + 1] per destination block: range in kt_idx
    const int32_t *kt_idx;          // source tiles (16 wide) holding at least one non-zero of the block
    const int32_t *bm_order;        // [Mp/128] destination blocks by decreasing number of source tiles (launch order)
    const int32_t *rq_beg, *rq_idx; // [nriv+1], river entries (index into DevCatchment::riv) per river, ascending HRU
    const int32_t *ro_beg, *ro_idx; // [nriv+1], HRUs with ipriv == r, ascending
};

// ---- resident (persistent, on-chip state) kernel ------------------------------------------------
struct __align__(8) RivChain {      // one term of qb(r): (((dtt*qbf(ia))*w_riv)*share)*ac   (dyna_dynamic_dist.f90:59-61)
    double w_riv, share, ac;
    int32_t ia;
    int32_t pad;
};

struct ResPlan {
    int32_t feasible;
    int32_t lean_ok;                // every HRU has ims_rz == ims_uz == 1 and ntt == 1: the branch-free step applies
    int32_t nac_pad;                // HRUs per member, padded to a power of two <= 512
    int32_t G;                      // member sub-groups side by side in one CTA (512 / nac_pad)
    int32_t Mc;                     // members per CTA = G * S
    int32_t n_groups;               // ceil(B / Mc), set per launch
    int32_t n_threads;              // 512 compute threads + 128 river threads
    int32_t variant;                // 1: one HRU column per compute thread (16 warps), 2: two columns (8 warps, 232 registers)
    int32_t river_team;             // lanes per summation chain in FAST mode (power of two <= 16)
    int32_t Tt;                     // forcing tile (timesteps) per TMA bulk copy
    int32_t nnz;                    // W entries
    int32_t nq;                     // non-zero entries of rivers(:, :)
    int32_t smem_bytes;
    int32_t off_qbf, off_ex, off_cst, off_par, off_w, off_src, off_ac, off_forc, off_bar, off_rqc, off_rq, off_ro;
    const int32_t *thread_hru;      // [nac_pad] HRU handled at each thread position (-1 = idle), sorted by W row length
    const int32_t *hru_pos;         // [nac] inverse of thread_hru
    const RivChain *rivq;           // river CSR for qb, per river ascending HRU
    const int32_t *rivq_beg;        // [nriv+1]
    const int32_t *rivo;            // HRUs by ipriv, ascending
    const int32_t *rivo_beg;        // [nriv+1]
    long long *dbg;                 // optional [grid][20 warps] busy-cycle counters (env DCP_RES_DEBUG=1), else null
};

// ---- resident kernel, second generation (lean arithmetic only; dcp_resident_v2.cu) ------------------
// Two independent teams per CTA, each running 2 member slots over all 512 HRU columns (4 columns per
// compute thread).  The HRU->HRU weights are stored in ELLPACK form per block of 32 thread positions
// (step-major, rows padded with zero weights to the block's longest row rounded up to 4), flattened per
// compute warp into one gather list (the thread's four columns one after the other, source tuple index
// packed into the low mantissa bits of the weight), and every compute thread publishes its HRU's
// complete contribution to its river's inflow, so the river warps only add up contiguous ranges.
struct Res2Plan {
    int32_t feasible;
    int32_t nac_pad;                // HRUs per member, padded to a power of two in [32, 512]
    int32_t G;                      // member sub-groups side by side in one team (512 / nac_pad)
    int32_t n_groups;               // team groups of G*2 members: ceil(B / (2G)), set per launch
    int32_t Tt;                     // forcing tile (timesteps) per TMA bulk copy
    int32_t ell_len;                // gather-list entries (padded, all four compute warps)
    int32_t team;                   // river lanes per (sub-group, river) sum (power of two <= 32)
    int32_t n_teams;                // number of such sums: G * nriv
    int32_t n_pass;                 // passes of the team's 64 river lanes over the sums
    int32_t phase_delay;            // cycles the second team waits at kernel start (0: none)
    int32_t smem_bytes, team_bytes;
    int32_t off_team;               // start of the team-private region; offsets below off_forc.. are relative to a team's base
    int32_t off_qbf, off_con, off_cst, off_par, off_forc, off_bar, off_job, off_acc;
    int32_t off_ellw, off_coli, off_cold, off_rbeg, off_eq0, off_fmtab;   // shared, absolute
    int32_t wide;                   // 1: plan of the wide build (one team, 1024 columns)
    const double *ell_w;            // [ell_len] w * ac(src)   (dyna_dynamic_dist.f90:50-51), 0 in padding; the low 9
                                    //           mantissa bits hold the source's tuple index (0..511)
    const int4 *col_i;              // [4][128] x: first gather-list entry of the column | K << 20;  y: ippt | ipet << 8 | g << 16 | idle << 31;
                                    //          z: tuple index | sub-group base << 16;  w: par offset | contribution slot << 16
    const double2 *col_d;           // [4][128][2] {ac, 1/ac}, {dtt*w_riv*share*ac, cos(mslp)}
    const int32_t *col_ia;          // [4][128] HRU of the column (-1 idle)
    const double *eq0;              // wide build: [4][256] exp(-st) of the column
    const double *fm_tab;           // narrow build: [64 + 256] tables of fm_exp_tab / fm_log_tab (dcp_fastmath.cuh)
    const int32_t *rbeg;            // [nriv+1] contribution-slot ranges per river
    // routing inside the kernel: jobs of ONE member (gauge i x node u in {i} + upstream tree of i, gauge-major, own node first)
    int32_t jobs_per_member;
    const int32_t *job_node;        // [jobs_per_member] node u of the job
    const int32_t *job_gauge;       // [jobs_per_member] gauge i of the job
    const int32_t *gauge_job_beg;   // [nriv+1] jobs of gauge i
    const double *inv_frac;         // [nriv] 1 / frac_sub_area of the gauge's node
    const double *obs_eps;          // [nriv] floor of the logarithmic measure (mean observed flow / 100)
    const double *log_obs;          // [nriv x nstep_obs] ln(max(qobs, eps)) (0 where the observation is missing), or null
    const double *pet_pos;          // [nge x nstep_pad] max(pe_step, 0): the root zone only acts on ep > 0 (dyna_root_zone1.f90:55)
    long long *dbg;                 // optional busy-cycle counters [grid][12 warps]
};
