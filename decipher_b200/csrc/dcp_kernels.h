// dcp_kernels.h — launch wrappers shared between the kernel TUs and the C-ABI host code.
#pragma once
#include "dcp_device.cuh"

#define DCP_STREAM_THREADS 128
#define DCP_RES_S 4                  // member slots per compute thread of the resident kernel
#define DCP_RES_LG 2                 // member slots interleaved per basic block in the lean step
#define DCP_RES_LG2 2                // same for the two-columns-per-thread variant
#ifndef DCP_RES2_GB
#define DCP_RES2_GB 1                // gather batch of the second-generation kernel: W rows are padded to a multiple of it
#endif
#define DCP_RES_COMPUTE 512
#define DCP_RES_MAXPAIRS 192         // (members per CTA) x nriv the river warps can carry (3 summing warps x 4 passes x 16 pairs)
#define DCP_RES_TT 64                // forcing tile length (timesteps)
#define DCP_RT_SLACK_STEPS 129        // = DCP_RT_SLACK of dcp_route_fused.cuh: ring length >= histogram span + two hand-over tiles + 1
#define DCP_FORCING_PAD 64           // forcing arrays are zero-padded to a multiple of this many steps

void dcp_internal_set_error(const char *msg);       // sets the calling thread's dcp_last_error() text (dcp_multi.cu)
void dcp_launch_min_chv(const double *mcpar, int npt, int64_t n_member, double *out, cudaStream_t s);
void dcp_launch_init(const DevCatchment &C, const DevBatch &W, double *mcpar, bool check, cudaStream_t s);
cudaError_t dcp_launch_physics_streamed(const DevCatchment &C, const DevBatch &W, int t0, int t1, bool fast,
                                        bool check, bool aet, cudaStream_t s);
cudaError_t dcp_launch_route(const DevCatchment &C, const DevBatch &W, int t0, int t1, double *q_out, cudaStream_t s);
void dcp_launch_route_prepare(const DevCatchment &C, const DevBatch &W, const double *v, const double *inflow,
                              cudaStream_t s);
void dcp_launch_route_status(const DevCatchment &C, const DevBatch &W, int32_t *status, cudaStream_t s);
void dcp_launch_finalize(const DevCatchment &C, const DevBatch &W, double *perf, double *totals, double *diag,
                         int32_t *status, double *sig, cudaStream_t s);
int dcp_select_blocks(int64_t n);
void dcp_launch_select_count(const double *perf, int64_t n, int nriv, int measure, int gauge, double thr, int hib, int32_t *block_count, cudaStream_t s);
void dcp_launch_select_scatter(const double *perf, int64_t n, int nriv, int measure, int gauge, double thr, int hib, const int64_t *block_off,
                               int64_t *index_out, double *dist_out, cudaStream_t s);
void dcp_launch_or_status(int32_t *status, const int32_t *extra, int32_t mask, int64_t n, cudaStream_t s);
void dcp_launch_fp64_peak(double *out, int iters, int blocks, cudaStream_t s);
cudaError_t dcp_launch_physics_resident(const DevCatchment &C, const DevBatch &W, const ResPlan &P, bool fast,
                                        bool aet, int grid, cudaStream_t s);
cudaError_t dcp_launch_physics_resident_v2(const DevCatchment &C, const DevBatch &W, const Res2Plan &P, bool aet,
                                           int grid, cudaStream_t s);
cudaError_t dcp_launch_physics_resident_v2w(const DevCatchment &C, const DevBatch &W, const Res2Plan &P, bool aet,
                                            int grid, cudaStream_t s);
cudaError_t dcp_launch_physics_stepwise(const DevCatchment &C, const DevBatch &W, const DensePlan &D, int t0, int t1,
                                        bool fast, bool aet, bool dense, int n_sm, cudaStream_t s);
void dcp_launch_debug_fastmath(int op, int64_t n, const double *x, double *y, const double *tab, cudaStream_t s);
