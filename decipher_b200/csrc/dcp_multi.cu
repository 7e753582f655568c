// dcp_multi.cu — the Monte-Carlo loop over ALL GPUs of one box behind the C ABI (include/decipher_b200.h:
// dcp_comm_init / dcp_run_ensemble_multi / dcp_gather_perf).
//
// The reference's loop `do i_mc = 1, numsim` (RRMODEL/dyna_main.f90:166-206) is serial; members never interact
// (all state is re-initialised per member: dyna_initialise_run.f90:27-36, dyna_init_satzone.f90:194-222,
// DTA/dta_route_processing.f90:279-292).  Here one host thread per rank drives one device: rank r runs the
// contiguous block [r*M/G + min(r, M%G), ...) of the ONE parameter table the caller sampled (sequential RANMAR), so
// results do not depend on the number of ranks.  No collective touches the data path; the only exchange is ONE
// gather of the per-member measures (+ status words, init diagnostics) to rank 0 at the end of the run:
//   * ranks on distinct devices : NCCL point-to-point inside one group (ncclSend on the ranks, ncclRecv on rank 0),
//                                 communicators from ncclCommInitAll; NCCL is loaded at run time (libnccl.so.2), so the
//                                 single-GPU entry points carry no NCCL dependency;
//   * ranks sharing a device    : (testing several ranks on one GPU) device-to-device copies into rank 0's buffer.
// Receive buffers live in the communicator and are reused by every run.
#include "dcp_kernels.h"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

int mfail(int code, const char *fmt, ...) {      // message goes to the calling thread's dcp_last_error()
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    dcp_internal_set_error(buf);
    return code;
}

struct NcclApi {                    // resolved once from libnccl.so.2
    void *h = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string &why) {
        if (h) return true;
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { why = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
        auto sym = [&](const char *n) { void *p = dlsym(h, n); if (!p) why = std::string("missing NCCL symbol ") + n; return p; };
        CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
        Send = (decltype(Send))sym("ncclSend");
        Recv = (decltype(Recv))sym("ncclRecv");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
        return CommInitAll && CommDestroy && GroupStart && GroupEnd && Send && Recv && GetErrorString;
    }
};
NcclApi g_nccl;

void block_of(int64_t n, int rank, int world, int64_t &lo, int64_t &hi) {     // decipher_b200/distributed.py: shard_range
    const int64_t base = n / world, rem = n % world;
    lo = rank * base + std::min<int64_t>(rank, rem);
    hi = lo + base + (rank < rem ? 1 : 0);
}

}  // namespace

struct dcp_comm {
    int n_ranks = 0;
    std::vector<int> device;
    bool use_nccl = false;                  // all devices distinct and more than one rank
    std::vector<ncclComm_t> comm;           // [n_ranks] when use_nccl
    std::vector<cudaStream_t> stream;       // [n_ranks] gather streams
    // rank 0 receive buffers (device 0 of the communicator), grown on demand and reused
    void *root_buf = nullptr;
    size_t root_cap = 0;
    // per-rank send staging (perf | diag | status of the rank's block), grown on demand and reused
    std::vector<void *> send_buf;
    std::vector<size_t> send_cap;
};

extern "C" {

int dcp_comm_size(const dcp_comm *c) { return c ? c->n_ranks : 0; }
int dcp_comm_uses_nccl(const dcp_comm *c) { return c && c->use_nccl ? 1 : 0; }

int dcp_comm_destroy(dcp_comm *c) {
    if (!c) return DCP_OK;
    for (int r = 0; r < c->n_ranks; ++r) {
        cudaSetDevice(c->device[r]);
        if (c->use_nccl && r < (int)c->comm.size() && c->comm[r]) g_nccl.CommDestroy(c->comm[r]);
        if (r < (int)c->stream.size() && c->stream[r]) cudaStreamDestroy(c->stream[r]);
        if (r < (int)c->send_buf.size() && c->send_buf[r]) cudaFree(c->send_buf[r]);
    }
    if (c->root_buf) { cudaSetDevice(c->device[0]); cudaFree(c->root_buf); }
    delete c;
    return DCP_OK;
}

int dcp_comm_init(int n_ranks, const int *devices, dcp_comm **out) {
    if (!out || n_ranks < 1) return mfail(DCP_ERR_INVALID, "dcp_comm_init: n_ranks < 1 or null output");
    *out = nullptr;
    const int n_dev = dcp_device_count();
    if (n_dev < 1) return mfail(DCP_ERR_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
    dcp_comm *c = new dcp_comm();
    c->n_ranks = n_ranks;
    c->device.resize(n_ranks);
    bool distinct = true;
    for (int r = 0; r < n_ranks; ++r) {
        c->device[r] = devices ? devices[r] : r;
        if (c->device[r] < 0 || c->device[r] >= n_dev) {
            const int bad = c->device[r];
            delete c;
            return mfail(DCP_ERR_INVALID, "dcp_comm_init: rank %d wants device %d, %d visible", r, bad, n_dev);
        }
        for (int q = 0; q < r; ++q) if (c->device[q] == c->device[r]) distinct = false;
    }
    c->stream.assign(n_ranks, nullptr);
    c->send_buf.assign(n_ranks, nullptr);
    c->send_cap.assign(n_ranks, 0);
    for (int r = 0; r < n_ranks; ++r) {
        if (cudaSetDevice(c->device[r]) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream[r], cudaStreamNonBlocking) != cudaSuccess) {
            const int code = mfail(DCP_ERR_CUDA, "dcp_comm_init: stream on device %d: %s", c->device[r], cudaGetErrorString(cudaGetLastError()));
            dcp_comm_destroy(c);
            return code;
        }
    }
    c->use_nccl = distinct && n_ranks > 1;
    if (c->use_nccl) {
        std::string why;
        if (!g_nccl.load(why)) { dcp_comm_destroy(c); return mfail(DCP_ERR_UNSUPPORTED, "NCCL unavailable: %s", why.c_str()); }
        c->comm.assign(n_ranks, nullptr);
        const ncclResult_t rc = g_nccl.CommInitAll(c->comm.data(), n_ranks, c->device.data());
        if (rc != ncclSuccess) {
            c->use_nccl = false;
            dcp_comm_destroy(c);
            return mfail(DCP_ERR_CUDA, "ncclCommInitAll: %s", g_nccl.GetErrorString(rc));
        }
    }
    *out = c;
    return DCP_OK;
}

// The collective alone: every rank holds `bytes[r]` bytes at send[r] (device memory of rank r's device); on return they
// sit back to back, in rank order, in the communicator's root buffer (device of rank 0), whose address is *root_out.
// Called from ONE host thread; enqueues on the communicator's streams and waits for them.
int dcp_gather_bytes(dcp_comm *c, const void *const *send, const size_t *bytes, void **root_out) {
    if (!c || !send || !bytes || !root_out) return mfail(DCP_ERR_INVALID, "dcp_gather_bytes: null argument");
    size_t total = 0;
    for (int r = 0; r < c->n_ranks; ++r) total += bytes[r];
    cudaSetDevice(c->device[0]);
    if (total > c->root_cap) {
        if (c->root_buf) cudaFree(c->root_buf);
        c->root_buf = nullptr; c->root_cap = 0;
        if (cudaMalloc(&c->root_buf, std::max<size_t>(total, 256)) != cudaSuccess)
            return mfail(DCP_ERR_NOMEM, "gather buffer of %zu bytes: %s", total, cudaGetErrorString(cudaGetLastError()));
        c->root_cap = std::max<size_t>(total, 256);
    }
    std::vector<size_t> off(c->n_ranks, 0);
    for (int r = 1; r < c->n_ranks; ++r) off[r] = off[r - 1] + bytes[r - 1];
    if (c->use_nccl) {
        // one group: rank 0 posts every receive, the other ranks their send; rank 0's own block is a local copy
        ncclResult_t rc = g_nccl.GroupStart();
        for (int r = 1; r < c->n_ranks && rc == ncclSuccess; ++r) {
            if (bytes[r] == 0) continue;
            rc = g_nccl.Recv((char *)c->root_buf + off[r], bytes[r], ncclChar, r, c->comm[0], c->stream[0]);
            if (rc == ncclSuccess) rc = g_nccl.Send(send[r], bytes[r], ncclChar, 0, c->comm[r], c->stream[r]);
        }
        const ncclResult_t rc2 = g_nccl.GroupEnd();
        if (rc != ncclSuccess || rc2 != ncclSuccess)
            return mfail(DCP_ERR_CUDA, "NCCL gather: %s", g_nccl.GetErrorString(rc != ncclSuccess ? rc : rc2));
        cudaSetDevice(c->device[0]);
        if (bytes[0] && cudaMemcpyAsync(c->root_buf, send[0], bytes[0], cudaMemcpyDeviceToDevice, c->stream[0]) != cudaSuccess)
            return mfail(DCP_ERR_CUDA, "gather: local copy: %s", cudaGetErrorString(cudaGetLastError()));
    } else {
        for (int r = 0; r < c->n_ranks; ++r) {
            if (bytes[r] == 0) continue;
            cudaSetDevice(c->device[r]);
            const cudaError_t e = c->device[r] == c->device[0]
                ? cudaMemcpyAsync((char *)c->root_buf + off[r], send[r], bytes[r], cudaMemcpyDeviceToDevice, c->stream[r])
                : cudaMemcpyPeerAsync((char *)c->root_buf + off[r], c->device[0], send[r], c->device[r], bytes[r], c->stream[r]);
            if (e != cudaSuccess) return mfail(DCP_ERR_CUDA, "gather: device copy of rank %d: %s", r, cudaGetErrorString(e));
        }
    }
    for (int r = 0; r < c->n_ranks; ++r) {
        cudaSetDevice(c->device[r]);
        const cudaError_t e = cudaStreamSynchronize(c->stream[r]);
        if (e != cudaSuccess) return mfail(DCP_ERR_CUDA, "gather: rank %d: %s", r, cudaGetErrorString(e));
    }
    *root_out = c->root_buf;
    return DCP_OK;
}

int dcp_run_ensemble_multi(dcp_comm *c, const dcp_catchment_desc *desc, int64_t n_member, double *mcpar,
                           const dcp_run_opts *opts, double *q_out, double *perf_out, double *reach_totals,
                           double *init_diag, int32_t *status, dcp_multi_stats *stats) {
    if (!c || !desc || !mcpar) return mfail(DCP_ERR_INVALID, "dcp_run_ensemble_multi: null communicator, descriptor or mcpar");
    if (n_member < 1) return mfail(DCP_ERR_INVALID, "n_member < 1");
    dcp_run_opts o{};
    o.struct_size = sizeof(dcp_run_opts);
    if (opts) {
        if (opts->struct_size != (int)sizeof(dcp_run_opts)) return mfail(DCP_ERR_INVALID, "dcp_run_opts.struct_size mismatch");
        o = *opts;
    }
    if (o.flags & DCP_OPT_DEVICE_BUFFERS) return mfail(DCP_ERR_INVALID, "dcp_run_ensemble_multi takes host buffers");
    const int G = c->n_ranks, nriv = desc->nriv;
    const size_t npar = (size_t)desc->npt * 7;
    // Large outputs (flow series, reach totals) go straight from each device into the caller's arrays; then nothing
    // is left to gather.  The measures-only run keeps everything on the devices and gathers once at the end.
    const bool want_sig = (o.flags & DCP_OPT_SIGNATURES) != 0;
    if (want_sig && !o.sig_out) return mfail(DCP_ERR_INVALID, "DCP_OPT_SIGNATURES without dcp_run_opts.sig_out");
    const bool direct = q_out != nullptr || reach_totals != nullptr || want_sig;
    const size_t perf_b = perf_out ? (size_t)DCP_NMEASURE * nriv * 8 : 0, diag_b = init_diag ? 24 : 0, st_b = status ? 4 : 0;
    std::vector<int> rc(G, DCP_OK);
    std::vector<std::string> err(G);
    std::vector<dcp_run_stats> rstats(G);
    std::vector<double> wall_ms(G, 0.0);
    std::vector<const void *> send(G, nullptr);
    std::vector<size_t> send_bytes(G, 0);
    const auto t_all = std::chrono::steady_clock::now();

    auto worker = [&](int r) {
        int64_t lo, hi;
        block_of(n_member, r, G, lo, hi);
        const int64_t n = hi - lo;
        if (n == 0) return;
        const auto t0 = std::chrono::steady_clock::now();
        auto bail = [&](int code, const char *what, const char *msg) { rc[r] = code; err[r] = std::string(what) + ": " + msg; };
        if (cudaSetDevice(c->device[r]) != cudaSuccess) return bail(DCP_ERR_CUDA, "cudaSetDevice", cudaGetErrorString(cudaGetLastError()));
        dcp_catchment *cat = nullptr;
        int k = dcp_catchment_create(desc, &cat);
        if (k != DCP_OK) return bail(k, "dcp_catchment_create", dcp_last_error());
        if (direct) {
            dcp_run_opts orank = o;                      // this rank's block of the signature table
            if (want_sig) orank.sig_out = o.sig_out + (size_t)DCP_NSIG * nriv * lo;
            k = dcp_run_ensemble(cat, lo + 1, n, mcpar + npar * lo, &orank, q_out ? q_out + (size_t)nriv * desc->nstep * lo : nullptr,
                                 perf_out ? perf_out + (size_t)DCP_NMEASURE * nriv * lo : nullptr,
                                 reach_totals ? reach_totals + (size_t)3 * nriv * lo : nullptr,
                                 init_diag ? init_diag + 3 * lo : nullptr, status ? status + lo : nullptr);
            if (k != DCP_OK) bail(k, "dcp_run_ensemble", dcp_last_error());
        } else {
            // device-resident run: [perf | diag | status] of the block form the rank's send buffer, mcpar is staged beside it
            const size_t out_b = (perf_b + diag_b + st_b) * (size_t)n, need = out_b + 256 + npar * 8 * (size_t)n;
            if (need > c->send_cap[r]) {
                if (c->send_buf[r]) cudaFree(c->send_buf[r]);
                c->send_buf[r] = nullptr; c->send_cap[r] = 0;
                if (cudaMalloc(&c->send_buf[r], need) != cudaSuccess) { dcp_catchment_destroy(cat); return bail(DCP_ERR_NOMEM, "send buffer", cudaGetErrorString(cudaGetLastError())); }
                c->send_cap[r] = need;
            }
            char *b = (char *)c->send_buf[r];
            double *d_perf = perf_b ? (double *)b : nullptr;
            double *d_diag = diag_b ? (double *)(b + perf_b * n) : nullptr;
            int32_t *d_status = st_b ? (int32_t *)(b + (perf_b + diag_b) * n) : nullptr;
            double *d_mc = (double *)(b + ((out_b + 255) & ~(size_t)255));
            cudaError_t e = cudaMemcpy(d_mc, mcpar + npar * lo, npar * 8 * (size_t)n, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) {
                dcp_run_opts od = o;
                od.flags |= DCP_OPT_DEVICE_BUFFERS;
                k = dcp_run_ensemble(cat, lo + 1, n, d_mc, &od, nullptr, d_perf, nullptr, d_diag, d_status);
                if (k != DCP_OK) bail(k, "dcp_run_ensemble", dcp_last_error());
                else e = cudaMemcpy(mcpar + npar * lo, d_mc, npar * 8 * (size_t)n, cudaMemcpyDeviceToHost);    // srinit clip is in/out
            }
            if (e != cudaSuccess) bail(DCP_ERR_CUDA, "parameter copy", cudaGetErrorString(e));
            send[r] = b; send_bytes[r] = out_b;
        }
        dcp_get_run_stats(cat, &rstats[r]);
        dcp_catchment_destroy(cat);
        wall_ms[r] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    std::vector<std::thread> th;
    for (int r = 1; r < G; ++r) th.emplace_back(worker, r);
    worker(0);
    for (auto &t : th) t.join();
    for (int r = 0; r < G; ++r)
        if (rc[r] != DCP_OK) return mfail(rc[r], "rank %d (device %d): %s", r, c->device[r], err[r].c_str());

    double gather_ms = 0.0;
    if (!direct && (perf_b + diag_b + st_b) > 0) {
        const auto tg = std::chrono::steady_clock::now();
        void *root = nullptr;
        const int k = dcp_gather_bytes(c, send.data(), send_bytes.data(), &root);
        if (k != DCP_OK) return k;
        // rank 0 unpacks the blocks [perf | diag | status] into the caller's arrays
        cudaSetDevice(c->device[0]);
        size_t off = 0;
        for (int r = 0; r < G; ++r) {
            int64_t lo, hi;
            block_of(n_member, r, G, lo, hi);
            const size_t n = (size_t)(hi - lo);
            const char *b = (const char *)root + off;
            cudaError_t e = cudaSuccess;
            if (perf_b && e == cudaSuccess) e = cudaMemcpy(perf_out + (size_t)DCP_NMEASURE * nriv * lo, b, perf_b * n, cudaMemcpyDeviceToHost);
            if (diag_b && e == cudaSuccess) e = cudaMemcpy(init_diag + 3 * lo, b + perf_b * n, diag_b * n, cudaMemcpyDeviceToHost);
            if (st_b && e == cudaSuccess) e = cudaMemcpy(status + lo, b + (perf_b + diag_b) * n, st_b * n, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) return mfail(DCP_ERR_CUDA, "gather: copy to host: %s", cudaGetErrorString(e));
            off += send_bytes[r];
        }
        gather_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tg).count();
    }
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->n_ranks = G;
        stats->used_nccl = c->use_nccl && !direct;
        stats->ms_gather = gather_ms;
        stats->ms_wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_all).count();
        for (int r = 0; r < G && r < DCP_MAX_RANKS; ++r) {
            stats->ms_rank_wall[r] = wall_ms[r];
            stats->ms_rank_device[r] = rstats[r].ms_total;
            stats->ms_busy_max = std::max(stats->ms_busy_max, rstats[r].ms_total);
            stats->ms_busy_sum += rstats[r].ms_total;
        }
        stats->kernel_used = rstats[0].kernel_used;
    }
    return DCP_OK;
}

}  // extern "C"
