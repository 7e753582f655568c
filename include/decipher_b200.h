/*
 * decipher_b200.h — C ABI of the B200-native DECIPHeR Monte-Carlo rainfall-runoff path.
 *
 * This is the drop-in boundary.  The reference has no FFI layer; the seam it offers is the
 * Fortran call
 *
 *     call mainloop(all_pm_names, nac, nstep, num_rivers, acc, dt, dtt, dyna_hru, mcpar, ntt,
 *                   node_to_flow_mapping, num_par_types, pe_step, qobs_riv_step_start,
 *                   r_gau_step, rivers, route_riv, route_tdh, sum_ac_riv, wt)
 *
 * made once per ensemble member from the serial loop in RRMODEL/dyna_main.f90:166-206
 * (definition RRMODEL/dyna_main_loop.f90:9-28).  The functions below replace that loop body
 * for ALL members at once: the caller samples every parameter set first (genpar,
 * RRMODEL/dyna_genpar.f90:35-41), then makes ONE dcp_run_ensemble call.
 *
 * Conventions (chosen so that a Fortran `bind(C)` interface is 1:1):
 *   - plain C, no exceptions cross the boundary, int32_t / int64_t / double / opaque handles only;
 *   - every array is contiguous and laid out exactly as the Fortran reference stores it
 *     (column-major, first index fastest);
 *   - every index stored IN an array (itrans, ipar, ipriv, ippt, ipet, upstream tree, mapping)
 *     stays 1-based as in the reference files;
 *   - the library copies what it needs at create time; the caller may free its buffers after
 *     dcp_catchment_create returns;
 *   - return value 0 = DCP_OK, anything else is an error code and dcp_last_error() describes it;
 *   - there is no CPU fallback and no backend dispatch: without a CUDA device every compute
 *     entry point returns DCP_ERR_NO_DEVICE.
 */
#ifndef DECIPHER_B200_H
#define DECIPHER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCP_ABI_VERSION 4   /* 2: DCP_NMEASURE 2 -> 5 (layout of perf_out) and dcp_route_timeseries; 3: dcp_comm_* / dcp_run_ensemble_multi;
                               4: dcp_run_opts.sig_out (flow-duration signatures), dcp_select_behavioural */

/* ---- error codes ------------------------------------------------------------------------- */
#define DCP_OK               0
#define DCP_ERR_INVALID      1   /* bad argument / inconsistent descriptor                      */
#define DCP_ERR_NO_DEVICE    2   /* no CUDA device, or the device is not sm_100                 */
#define DCP_ERR_CUDA         3   /* a CUDA runtime call failed (message holds cudaGetErrorString) */
#define DCP_ERR_NOMEM        4   /* host or device allocation failed                            */
#define DCP_ERR_UNSUPPORTED  5   /* a reference feature this build does not implement           */

/* ---- per-member status word (dcp_run_ensemble: status[]) ---------------------------------
 * The reference STOPs the whole process on these conditions; the ensemble path records them
 * per member instead and keeps going.                                                         */
#define DCP_ST_SOLFLAG_MASK   0x3       /* init_satzone sol_flag 0/1/2 (dyna_init_satzone.f90:120-186) */
#define DCP_ST_RZ_BALANCE     0x4       /* root-zone store balance STOP (dyna_root_zone1.f90:73-77)   */
#define DCP_ST_TS_OUTSUM      0x8       /* output-sum STOP (dyna_results_ts.f90:88-93)                */
#define DCP_ST_TS_WBAL        0x10      /* water-balance STOP (dyna_results_ts.f90:95-101)            */
#define DCP_ST_NONFINITE      0x20      /* a routed flow of this member is inf/NaN                    */
#define DCP_ST_INIT_ITERCAP   0x40      /* init search hit DCP_INIT_MAX_ITER (guard, not in reference)*/
#define DCP_ST_HIST_DROPPED   0x80      /* hist_add_value dropped a cell (dta_route_processing.f90:378-382) */

#define DCP_INIT_MAX_ITER     100000000 /* cap on the while loop of dyna_init_satzone.f90:134-174     */

/* ---- run option flags -------------------------------------------------------------------- */
#define DCP_OPT_CHECK_BALANCE   0x1     /* evaluate the results_ts / root_zone1 STOP conditions: they are statements
                                           about the reference's own rounding, so the streamed kernel evaluates them
                                           in the reference's arithmetic -- DCP_KERNEL_AUTO runs the whole ensemble
                                           there, an explicit DCP_KERNEL_RESIDENT / _STEPWISE request keeps its kernel
                                           for every output and adds one streamed pass that contributes the flag bits */
#define DCP_OPT_DEVICE_BUFFERS  0x2     /* mcpar / outputs are DEVICE pointers (no H2D/D2H copies)    */
#define DCP_OPT_FAST_MATH       0x4     /* reciprocal-hoisted arithmetic (<=1e-9 rel.; default is the
                                           reference's operation order with IEEE division)            */

#define DCP_OPT_SIGNATURES      0x8     /* flow-duration-curve signatures into dcp_run_opts.sig_out (extension)    */

/* kernel selector (dcp_run_opts.kernel) */
#define DCP_KERNEL_AUTO      0
#define DCP_KERNEL_STREAMED  1          /* state streamed through L2/HBM, any nac                     */
#define DCP_KERNEL_RESIDENT  2          /* persistent kernel, state on chip for the whole simulation  */
#define DCP_KERNEL_STEPWISE  3          /* one timestep = redistribute (CSR SpMM, or an FP64 GEMM when the
                                           dynamic_dist matrix is dense) + physics over (HRU, member) +
                                           ordered river sums; any nac, no balance checks              */

/* performance measures per (gauge, member), in this order (definitions: DESIGN.md "Objective"; none exists upstream) */
#define DCP_NMEASURE 5
#define DCP_PERF_NSE    0   /* Nash-Sutcliffe efficiency                                  */
#define DCP_PERF_LOGNSE 1   /* NSE of ln(max(x, mean_obs/100))                            */
#define DCP_PERF_KGE    2   /* Kling-Gupta efficiency (2009 form: r, sigma ratio, mean ratio) */
#define DCP_PERF_PBIAS  3   /* 100 * sum(sim - obs) / sum(obs)                            */
#define DCP_PERF_RMSE   4   /* root mean square error, m/timestep                         */

/* Flow-duration-curve signatures per (gauge, member) (extension, DESIGN.md "Objective"; nothing upstream): computed on the
 * device from a streaming histogram of the routed flow (steps after t_warm) -- no flow series has to leave the chip.
 * Histogram bins are cut on the IEEE bits of the flow: 8 linear sub-bins per octave over [2^-40, 2^8) m/timestep
 * (+ one bin below, one above); quantiles interpolate linearly inside their bin.                                      */
#define DCP_NSIG 4
#define DCP_SIG_Q5      0   /* flow exceeded 5 % of the time (high flows)                  */
#define DCP_SIG_Q50     1   /* median flow                                                 */
#define DCP_SIG_Q95     2   /* flow exceeded 95 % of the time (low flows)                  */
#define DCP_SIG_SLOPE   3   /* slope of the FDC mid-segment: (ln Q33 - ln Q66) / 0.33      */
#define DCP_FDC_OCTAVE_LO (-40)
#define DCP_FDC_OCTAVES   48
#define DCP_FDC_BINS      (8 * DCP_FDC_OCTAVES + 2)

/*
 * Static description of one catchment: everything the reference holds after
 * Tread_dyna / Inputs / modelstruct_setup / read_routingdata have run
 * (RRMODEL/dyna_main.f90:95-160), flattened out of its derived types.
 */
typedef struct dcp_catchment_desc {
    int32_t struct_size;        /* sizeof(dcp_catchment_desc), for ABI checking                       */
    int32_t nac;                /* number of HRUs                 (dyna_tread_dyna.f90:47)            */
    int32_t nriv;               /* num_rivers                     (dyna_tread_dyna.f90:115)           */
    int32_t npt;                /* num_par_types                  (dyna_mc_setup.f90:84)              */
    int32_t nstep;              /* timesteps of rain/PET          (dyna_inputs.f90:166-169)           */
    int32_t ngp;                /* rain gauges / grid ids         (dyna_inputs.f90:126)               */
    int32_t nge;                /* PET gauges / grid ids          (dyna_inputs.f90:148)               */
    int32_t ntt;                /* inner sub-steps NTT            (dyna_mc_setup.f90:83)              */
    double  dt;                 /* timestep length                (dyna_inputs.f90:148)               */

    /* per-HRU static data, all [nac] (dyna_common_types.f90:22-44) */
    const double  *ac;          /* fractional area, after the renormalisation of dyna_tread_dyna.f90:186-190 */
    const double  *st;          /* mean topographic index                                            */
    const double  *mslp;        /* mean slope (radians)                                              */
    const int32_t *ippt;        /* 1-based rain column                                               */
    const int32_t *ipet;        /* 1-based PET column                                                */
    const int32_t *ipar;        /* 1-based parameter layer                                           */
    const int32_t *ipriv;       /* 1-based river the HRU's overland flow drains to                   */
    const int32_t *ims_rz;      /* root-zone structure id (only 1 exists, dyna_topmod.f90:146)       */
    const int32_t *ims_uz;      /* unsat-zone structure id (only 1 exists, dyna_topmod.f90:161)      */

    /* HRU->HRU flux weights as read by dyna_tread_dyna.f90:133-148 (after :173-184) */
    const int32_t *ntrans;      /* [nac]                                                             */
    const int32_t *itrans;      /* [sum(ntrans)]   1-based target HRUs, HRU after HRU                */
    const double  *wtrans;      /* [sum(ntrans+1)] weights, the river weight LAST for each HRU       */
    const double  *rivers;      /* [nac x nriv] column-major, share of HRU ia going to river iriv    */
    const double  *sum_ac_riv;  /* [nriv]  (dyna_tread_dyna.f90:166-167,192-194)                     */
    const double  *qobs_start;  /* [nriv]  qobs_riv_step_start (dyna_inputs.f90:84-113)              */

    /* river network (route_river_info_type, DTA/dta_route_processing.f90:10-27) */
    int32_t nnode;              /* size(node_list)                                                   */
    int32_t ncell;              /* rows of river_data                                                */
    const int32_t *node_gauge_id;   /* [nnode]                                                       */
    const int32_t *node_type;       /* [nnode] 1 sea, 2 gauge, 3 river                               */
    const int32_t *uptree_count;    /* [nnode] upstream_tree_count                                   */
    const int32_t *uptree_index;    /* [sum(uptree_count)] 1-based node indexes in the order produced
                                       by riv_node_full_upstream (DTA/dta_riv_tree_node.f90:80-110)  */
    const double  *frac_sub_area;   /* [nnode] (dyna_read_routingdata.f90:52-89)                     */
    const double  *river_data;      /* [ncell x 8] column-major: riv_id, area, dist, section_dist,
                                       elevation, slope, ds_dist, ds_elevation                       */
    const int32_t *node_to_flow_mapping; /* [nriv] 1-based (dyna_read_routingdata.f90:100-102)       */

    /* forcing (dyna_inputs.f90:117-160), gauge index fastest */
    const double  *rain;        /* r_gau_step [ngp x nstep]                                          */
    const double  *pet;         /* pe_step    [nge x nstep]                                          */

    /* observed flow, optional (NULL: no performance measures).  The reference reads and drops it
       (dyna_inputs.f90:31,78,90); it is used only by the NSE / log-NSE extension. */
    const double  *qobs;        /* [nriv x nstep_obs]                                                */
    int32_t nstep_obs;
    int32_t t_warm;             /* steps 1..t_warm are excluded from the performance measures        */
    double  flow_err;           /* missing-data code of the flow file (dyna_inputs.f90:57)           */
} dcp_catchment_desc;

typedef struct dcp_run_opts {
    int32_t struct_size;        /* sizeof(dcp_run_opts)                                              */
    int32_t flags;              /* DCP_OPT_*                                                         */
    int32_t kernel;             /* DCP_KERNEL_*                                                      */
    int32_t members_per_batch;  /* 0 = library chooses                                               */
    int32_t steps_per_tile;     /* 0 = library chooses (time tile of the flow ring buffer)           */
    int32_t reserved;
    double *sig_out;            /* DCP_OPT_SIGNATURES: [DCP_NSIG x nriv x n_member] (host, or device with
                                   DCP_OPT_DEVICE_BUFFERS); ignored otherwise                          */
} dcp_run_opts;

/* timing of the last dcp_run_ensemble call on this handle, CUDA events on the library's stream */
typedef struct dcp_run_stats {
    double ms_total;            /* first kernel start -> last kernel end                              */
    double ms_init;             /* init_satzone + build_hist kernel                                   */
    double ms_physics;          /* time-loop kernel(s)                                                */
    double ms_route;            /* routing / objective kernel(s)                                      */
    double ms_h2d;              /* host->device copies inside the call                                */
    double ms_d2h;              /* device->host copies inside the call                                */
    int64_t n_physics_launches;
    int64_t n_launches;         /* all kernels launched by the call                                   */
    int32_t kernel_used;        /* DCP_KERNEL_STREAMED, _RESIDENT or _STEPWISE                        */
    int32_t hist_len;           /* T of route_processing_build_hist for the slowest member            */
} dcp_run_stats;

typedef struct dcp_catchment dcp_catchment;   /* opaque; one handle per (host thread, GPU)            */

/* Device management.  dcp_device_count() returns 0 without a usable GPU (never an error). */
int  dcp_device_count(void);
int  dcp_set_device(int device);
int  dcp_abi_version(void);
const char *dcp_last_error(void);            /* thread-local, valid until the next call              */

/* Replaces the state the reference builds in RRMODEL/dyna_main.f90:95-160 and keeps alive
   across the MC loop.  Validates the descriptor and copies it to the current device. */
int  dcp_catchment_create(const dcp_catchment_desc *desc, dcp_catchment **out);
int  dcp_catchment_destroy(dcp_catchment *c);

/*
 * Replaces the body of `do i_mc = 1, numsim` (RRMODEL/dyna_main.f90:166-206) for members
 * first_member .. first_member+n_member-1 (first_member is informational: the caller passes
 * the parameter sets of exactly these members).
 *
 *   mcpar         [npt x 7 x n_member]  in/out — column order szm, lnto, srmax, srinit, chv, td, smax
 *                 (dyna_param_setup.f90:58-70); srinit is clipped to srmax on return exactly as
 *                 dyna_init_satzone.f90:203-206 rewrites mcpar(ipar,4).
 *   q_out         [nriv x nstep x n_member] or NULL — routed flow q(iriv,it) in m/timestep
 *                 (dyna_river.f90:57-59), the numbers write_output prints to the .flow file.
 *   perf_out      [DCP_NMEASURE x nriv x n_member] or NULL — NSE, log-NSE, KGE, PBIAS, RMSE (extension).
 *   reach_totals  [3 x nriv x n_member] or NULL — precip, PET, actual-ET totals per reach in mm
 *                 (dyna_write_output.f90:61-77).
 *   init_diag     [3 x n_member] or NULL — mean_q_ac_weighted, qb_tmp, mean_q_ac_weighted_guess
 *                 (m/timestep; dyna_init_satzone.f90:188-190 prints them x1000).
 *   status        [n_member] or NULL — DCP_ST_* bit flags.
 *
 * Synchronous: all outputs are complete on return.
 */
int  dcp_run_ensemble(dcp_catchment *c, int64_t first_member, int64_t n_member,
                      double *mcpar, const dcp_run_opts *opts,
                      double *q_out, double *perf_out, double *reach_totals,
                      double *init_diag, int32_t *status);

/*
 * Standalone router: replaces route_processing_build_hist + route_process_run_timeseries
 * (DTA/dta_route_processing.f90:154-354, 583-630) as driven by DTA/route_run_standalone.f90:118-189,
 * for n_series independent (velocity, inflow series) pairs over the river network of `c`.
 *
 *   velocity  [n_series]                  tdh%v_param(1): metres travelled per timestep
 *                                         (calculate_cell_delay, dta_route_processing.f90:419-442, V_MODE_CONSTANT)
 *   inflow    [nriv x nstep x n_series]   timeseries_flow_input: column i feeds node node_to_flow_mapping(i)
 *   routed    [nriv x nstep x n_series]   timeseries_flow_output: the plain routed sums (NOT divided by
 *                                         frac_sub_area; that division is dyna_river.f90:57-59)
 *   status    [n_series] or NULL          DCP_ST_HIST_DROPPED / DCP_ST_NONFINITE
 *   flags     0 or DCP_OPT_DEVICE_BUFFERS
 */
int  dcp_route_timeseries(dcp_catchment *c, int64_t n_series, const double *velocity, const double *inflow,
                          double *routed, int32_t *status, int32_t flags);

int  dcp_get_run_stats(const dcp_catchment *c, dcp_run_stats *out);

/*
 * GLUE-style behavioural filtering (extension): members whose measure `measure` (DCP_PERF_*) passes
 * `threshold` at gauge `gauge` (0-based; -1: at EVERY gauge) are behavioural.  higher_is_better != 0: pass = value >=
 * threshold (NSE, log-NSE, KGE), else value <= threshold (RMSE; PBIAS is compared on its absolute value).  NaN never passes.
 *   perf        [DCP_NMEASURE x nriv x n_member], host -- or device with flags = DCP_OPT_DEVICE_BUFFERS
 *   index_out   [n_member] host: the behavioural members (0-based), ascending
 *   weight_out  [n_member] host or NULL: likelihood weights |value - threshold| (worst gauge when gauge = -1), normalised
 *               to sum 1 over the behavioural members (all equal when every distance is 0)
 * The mask and the order-preserving compaction run on the device.                                                  */
int  dcp_select_behavioural(const double *perf, int64_t n_member, int32_t nriv, int32_t measure, int32_t gauge,
                            double threshold, int32_t higher_is_better, int64_t *index_out, double *weight_out,
                            int64_t *n_selected, int32_t flags);

/* ---- all GPUs of one box (dcp_comm_init / dcp_gather_perf) ----------------------------------------
 * One process, one host thread per rank.  Rank r owns device devices[r] (NULL: rank r -> device r) and runs the
 * contiguous member block [r*M/G + min(r, M mod G), ...) of the caller's parameter table, so the results do not
 * depend on the number of ranks.  The data path has no collective; the only exchange is ONE gather of the per-member
 * measures, init diagnostics and status words to rank 0 at the end of the run -- NCCL point-to-point over NVLink when
 * every rank has its own device (libnccl.so.2 is loaded at run time), plain device copies when ranks share a device
 * (several ranks on one GPU: testing).                                                                           */
#define DCP_MAX_RANKS 16
typedef struct dcp_comm dcp_comm;
typedef struct dcp_multi_stats {
    int32_t n_ranks;
    int32_t used_nccl;          /* 1: the gather went through NCCL                                              */
    int32_t kernel_used;        /* of rank 0                                                                     */
    int32_t reserved;
    double ms_wall;             /* host wall clock of the whole call                                             */
    double ms_gather;           /* the end-of-run gather + copy to the caller's arrays                           */
    double ms_busy_max;         /* slowest rank, CUDA events (dcp_run_stats.ms_total)                            */
    double ms_busy_sum;
    double ms_rank_device[DCP_MAX_RANKS];   /* per rank: CUDA-event time of its dcp_run_ensemble                 */
    double ms_rank_wall[DCP_MAX_RANKS];     /* per rank: host wall clock incl. catchment upload                  */
} dcp_multi_stats;

int  dcp_comm_init(int n_ranks, const int *devices, dcp_comm **out);
int  dcp_comm_destroy(dcp_comm *c);
int  dcp_comm_size(const dcp_comm *c);
int  dcp_comm_uses_nccl(const dcp_comm *c);
/* Replaces the loop `do i_mc = 1, numsim` (RRMODEL/dyna_main.f90:166-206) over every rank of the communicator.
 * Arguments as dcp_run_ensemble (HOST buffers); when q_out or reach_totals is requested each device writes its
 * block straight into the caller's arrays and nothing is gathered.                                              */
int  dcp_run_ensemble_multi(dcp_comm *c, const dcp_catchment_desc *desc, int64_t n_member, double *mcpar,
                            const dcp_run_opts *opts, double *q_out, double *perf_out, double *reach_totals,
                            double *init_diag, int32_t *status, dcp_multi_stats *stats);
/* The collective alone: bytes[r] bytes at send[r] (device memory of rank r) -> back to back, in rank order, in the
 * communicator's reusable root buffer on rank 0's device (*root_out).                                           */
int  dcp_gather_bytes(dcp_comm *c, const void *const *send, const size_t *bytes, void **root_out);

/* Standalone FP64 pipe micro-benchmark (dependency-free DFMA chains) used as the roofline
   denominator; returns TFLOP/s (2 flops per DFMA) in *tflops. */
int  dcp_measure_fp64_peak(double *tflops, double *ms);

/* Test hook: evaluates the FAST policy's branch-free primitives on the device
   (op 0: exp, 1: log, 2: reciprocal, 3: table-driven exp, 4: table-driven log) for host arrays x -> y; used by the accuracy tests. */
int  dcp_debug_fastmath(int op, int64_t n, const double *x, double *y);

#ifdef __cplusplus
}
#endif
#endif /* DECIPHER_B200_H */
