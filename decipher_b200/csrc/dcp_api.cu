// dcp_api.cu — host side of the C ABI (include/decipher_b200.h).
//
// dcp_catchment_create flattens the descriptor into device structures once (the state the reference
// keeps alive across its MC loop, RRMODEL/dyna_main.f90:95-160); dcp_run_ensemble replaces the loop
// body for all members: batches of members, each batch = init kernel, then time tiles of
// {physics, routing convolution, gauge summation + objective}, then the finalize kernel.
// No CPU compute path exists here: without a device every entry point fails.
#include "dcp_kernels.h"
#include "dcp_fastmath.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <limits>
#include <string>
#include <functional>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return fail(e_ == cudaErrorMemoryAllocation ? DCP_ERR_NOMEM : DCP_ERR_CUDA,         \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct DevBuf {                     // growable device allocation
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }         // early returns (CUDA_TRY, fail) must not leak the allocation
};

template <class T>
cudaError_t upload(const std::vector<T> &v, const T **dst, std::vector<void *> &owned) {
    void *p = nullptr;
    const size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return e;
    owned.push_back(p);
    if (!v.empty()) e = cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    *dst = (const T *)p;
    return e;
}

}  // namespace

void dcp_internal_set_error(const char *msg) { g_err = msg ? msg : ""; }

struct dcp_catchment {
    int device = 0;
    DevCatchment C{};
    std::vector<void *> owned;          // static device arrays
    DevBuf ws;                          // per-batch workspace (one slab)
    DevBuf io;                          // staging for host<->device outputs
    cudaStream_t stream = nullptr;
    // host copies needed at run time
    std::vector<double> node_reach_dist, node_down_dist;
    const double *ones_riv = nullptr;   // [nriv] 1.0 (standalone router: no division by frac_sub_area)
    bool has_qobs = false;
    ResPlan plan{};                     // resident-kernel plan (plan.feasible == 0: streamed kernel only)
    Res2Plan plan2{};                   // second-generation resident kernel (lean arithmetic, own-river catchments)
    Res2Plan plan2w{};                  // its wide build: 513 .. 1024 HRUs (FAST policy only)
    DensePlan dplan{};                  // stepwise path: ordered river lists, dense W when it pays
    int n_sm = 148;
    size_t smem_optin = 0;
    std::string plan_why;               // why the resident kernel is not available
    dcp_run_stats stats{};
};

namespace {

// Plan of the second-generation resident kernel: ELL schedule of the weight matrix, column tables, river slots.
void build_plan2(dcp_catchment *c, const dcp_catchment_desc *d, const std::vector<HruStatic> &hru,
                 const std::vector<WEntry> &w, const std::vector<int32_t> &thread_hru, const std::vector<int32_t> &hru_pos,
                 int nac_pad, const std::vector<double> &pet, const std::vector<double> &frac, const std::vector<double> &qobs,
                 const std::vector<ObsStat> &obs, bool wide, cudaError_t &e) {
    // narrow build: 2 teams x 128 compute threads x 4 columns = 512 columns per member slot; wide build (dcp_resident_v2w.cu):
    // 1 team x 256 compute threads = 1024 columns
    const int nac = d->nac, nriv = d->nriv, npt = d->npt, S = 2, TEAMS = wide ? 1 : 2, TC = 256 / TEAMS, H = 4, COLS = H * TC;
    const int TR = 128 / TEAMS, NW = TC / 32, PARW = wide ? 4 : 3, COLW = wide ? 2 : 3;
    const DevCatchment &C = c->C;
    Res2Plan &Q = wide ? c->plan2w : c->plan2;
    Q.wide = wide ? 1 : 0;
    // every river share of an HRU must sit on its own river ipriv (contribution sums)
    std::vector<double> cq(nac, 0.0);
    for (int ia = 0; ia < nac; ++ia)
        for (int r = 0; r < nriv; ++r) {
            const double sh = d->rivers[(size_t)r * nac + ia];
            if (sh == 0.0) continue;
            if (r != hru[ia].ipriv) return;
            cq[ia] = C.dtt * hru[ia].w_riv * sh * hru[ia].ac;
        }
    const int G = COLS / nac_pad, npb = nac_pad / 32;
    if (d->ngp > 255 || d->nge > 255 || G * S * npt * PARW + npt * PARW > 65535) return;
    Q.nac_pad = nac_pad; Q.G = G; Q.Tt = DCP_RES_TT;
    Q.phase_delay = getenv("DCP_RES_PHASE") ? atoi(getenv("DCP_RES_PHASE")) : 0;

    // ---- ELL schedule per block of 32 positions.  Inside a quarter-warp (8 consecutive positions) one LDS.128
    // serves the lanes whose source tuples fall in different 16-byte bank classes (position mod 8), so the
    // entries of each row are ordered step by step to keep the classes of a quarter distinct.  (The order of the
    // terms of a row is free in the lean policy; the exact policy keeps the reference's ascending-source order.)
    std::vector<int32_t> ell_base(npb), ell_K(npb);
    std::vector<double> ell_w;
    std::vector<uint16_t> ell_src;
    const bool schedule = !getenv("DCP_RES_NO_ELLSCHED");
    int zero_tuple[8] = {-1, -1, -1, -1, -1, -1, -1, -1};            // an idle thread position per bank class
    for (int pos = 0; pos < nac_pad; ++pos) if (thread_hru[pos] < 0 && zero_tuple[pos & 7] < 0) zero_tuple[pos & 7] = pos;
    for (int pb = 0; pb < npb; ++pb) {
        int K = 0;
        for (int l = 0; l < 32; ++l) {
            const int hh = thread_hru[pb * 32 + l];
            if (hh >= 0) K = std::max(K, hru[hh].row_end - hru[hh].row_beg);
        }
        K = (K + DCP_RES2_GB - 1) / DCP_RES2_GB * DCP_RES2_GB;
        ell_base[pb] = (int)ell_w.size(); ell_K[pb] = K;
        const size_t base = ell_w.size();
        ell_w.resize(base + (size_t)K * 32, 0.0);
        ell_src.resize(base + (size_t)K * 32, 0);
        for (int q = 0; q < 4; ++q) {                                   // quarter-warp
            std::vector<std::vector<int>> rem(8);                       // remaining entry ids per row
            for (int l = 0; l < 8; ++l) {
                const int hh = thread_hru[pb * 32 + q * 8 + l];
                if (hh >= 0) for (int x = hru[hh].row_beg; x < hru[hh].row_end; ++x) rem[l].push_back(x);
            }
            for (int k = 0; k < K; ++k) {
                // One LDS.128 of a quarter-warp costs as many wavefronts as its most loaded 16-byte bank class.  Per step:
                // a maximum bipartite matching lanes -> classes (Kuhn's augmenting paths, rows without spare steps first)
                // over the classes each row still has entries in; a row that stays unmatched waits if it has spare steps,
                // and every waiting or finished row reads a zero-weight entry on a tuple of a still unused class.
                const int left = K - k;                                  // steps left including this one
                int owner[8];                                            // class -> lane
                int cls_of[8];                                           // lane -> class
                for (int c = 0; c < 8; ++c) { owner[c] = -1; cls_of[c] = -1; }
                bool has[8][8];
                for (int l = 0; l < 8; ++l) {
                    for (int c = 0; c < 8; ++c) has[l][c] = false;
                    for (int x : rem[l]) has[l][hru_pos[w[x].src] & 7] = true;
                }
                int order[8] = {0, 1, 2, 3, 4, 5, 6, 7};
                std::stable_sort(order, order + 8, [&](int a, int b) { return left - (int)rem[a].size() < left - (int)rem[b].size(); });
                std::function<bool(int, bool *)> augment = [&](int l, bool *seen) {
                    for (int c = 0; c < 8; ++c) {
                        if (!has[l][c] || seen[c]) continue;
                        seen[c] = true;
                        if (owner[c] < 0 || augment(owner[c], seen)) { owner[c] = l; cls_of[l] = c; return true; }
                    }
                    return false;
                };
                if (schedule)
                    for (int oi = 0; oi < 8; ++oi) {
                        const int l = order[oi];
                        if (rem[l].empty()) continue;
                        bool seen[8] = {false, false, false, false, false, false, false, false};
                        augment(l, seen);
                    }
                int load[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (int c = 0; c < 8; ++c) if (owner[c] >= 0) load[c] = 1;
                for (int oi = 0; oi < 8; ++oi) {                        // rows that could not get a class of their own
                    const int l = order[oi];
                    if (rem[l].empty() || cls_of[l] >= 0) continue;
                    if (schedule && (int)rem[l].size() < left) continue; // spare steps: wait
                    int best = -1;                                       // must advance: the least loaded class it has
                    for (int c = 0; c < 8; ++c) if (has[l][c] && (best < 0 || load[c] < load[best])) best = c;
                    if (!schedule) best = hru_pos[w[rem[l][0]].src] & 7;
                    cls_of[l] = best; load[best]++;
                }
                for (int l = 0; l < 8; ++l) {
                    const int pos = pb * 32 + q * 8 + l;
                    const size_t o = base + (size_t)k * 32 + q * 8 + l;
                    if (cls_of[l] >= 0) {
                        size_t pick = 0;
                        while ((hru_pos[w[rem[l][pick]].src] & 7) != cls_of[l]) ++pick;
                        const WEntry &we = w[rem[l][pick]];
                        // wide build: 1/ac of the DESTINATION is folded in too (qin is only ever used as qin / ac)
                        ell_w[o] = wide ? (we.w * we.ac_src) / hru[thread_hru[pos]].ac : we.w * we.ac_src;
                        ell_src[o] = (uint16_t)hru_pos[we.src];
                        rem[l].erase(rem[l].begin() + pick);
                    } else {
                        // zero weight on an idle position's tuple (always 0.0, so nothing non-finite can leak from another
                        // HRU) of the least loaded class; without a suitable idle position, on the row's own tuple
                        int best = -1;
                        for (int c = 0; c < 8; ++c) if (zero_tuple[c] >= 0 && (best < 0 || load[c] < load[best])) best = c;
                        if (best >= 0 && load[best] > load[pos & 7]) best = -1;
                        const int src = best >= 0 ? zero_tuple[best] : pos;
                        load[src & 7]++;
                        ell_w[o] = 0.0; ell_src[o] = (uint16_t)src;
                    }
                }
            }
        }
    }
    // ---- contribution slots: river-major, ascending HRU inside a river; idle positions after the last river ----
    // The order inside a river's range is free (the river warps add the whole range), so the slots are handed out per
    // quarter-warp of thread positions with distinct 16-byte bank classes (slot mod 8) wherever the ranges allow: the
    // compute threads' STS.128 into the slot array is then close to conflict-free (it was 10 wavefronts for 4).
    std::vector<int32_t> rank(nac_pad, 0), rbeg(nriv + 1, 0);
    int slot = 0;
    for (int r = 0; r < nriv; ++r) {
        for (int ia = 0; ia < nac; ++ia) if (hru[ia].ipriv == r) ++slot;
        rbeg[r + 1] = slot;
    }
    {
        const int n_rng = nriv + 1;                                   // last range: idle positions
        std::vector<std::vector<std::vector<int>>> free_s(n_rng, std::vector<std::vector<int>>(8));
        for (int r = 0; r < nriv; ++r) for (int x = rbeg[r + 1] - 1; x >= rbeg[r]; --x) free_s[r][x & 7].push_back(x);
        for (int x = nac_pad - 1; x >= slot; --x) free_s[nriv][x & 7].push_back(x);
        for (int q0 = 0; q0 < nac_pad; q0 += 8) {
            bool used[8] = {false, false, false, false, false, false, false, false};
            for (int pos = q0; pos < q0 + 8; ++pos) {
                const int ia = thread_hru[pos], r = ia >= 0 ? hru[ia].ipriv : nriv;
                int pick = -1;
                for (int c = 0; c < 8; ++c)                           // an unused class with the most free slots
                    if (!used[c] && !free_s[r][c].empty() && (pick < 0 || free_s[r][c].size() > free_s[r][pick].size())) pick = c;
                if (pick < 0)
                    for (int c = 0; c < 8; ++c)
                        if (!free_s[r][c].empty() && (pick < 0 || free_s[r][c].size() > free_s[r][pick].size())) pick = c;
                rank[pos] = free_s[r][pick].back();
                free_s[r][pick].pop_back();
                used[pick] = true;
            }
        }
    }

    // ---- the COLS/32 blocks of 32 columns are dealt to the compute warps of a team longest-first ----
    const int NBLK = COLS / 32;
    std::vector<int> blocks(NBLK), load(NW, 0), cnt(NW, 0);
    for (int i = 0; i < NBLK; ++i) blocks[i] = i;
    std::stable_sort(blocks.begin(), blocks.end(), [&](int a, int b) { return ell_K[a % npb] > ell_K[b % npb]; });
    std::vector<std::vector<int>> slot_blk(NW, std::vector<int>(H, 0));
    for (int bi = 0; bi < NBLK; ++bi) {
        const int blk = blocks[bi];
        int wsel = -1;
        for (int wi = 0; wi < NW; ++wi) if (cnt[wi] < H && (wsel < 0 || load[wi] < load[wsel])) wsel = wi;
        slot_blk[wsel][cnt[wsel]++] = blk;
        load[wsel] += ell_K[blk % npb];
    }
    // ---- flat gather list per compute warp: the four columns of a thread one after the other, step-major over the
    // lanes, one 8-byte word per entry = the weight with the byte offset of the source tuple (index << 4, index < COLS)
    // in mantissa bits 4..12 (4..13 in the wide build; weight rounded to the nearest multiple of 2^13 / 2^14 ulp: relative
    // change <= 2^-40 / 2^-39).  One spare batch of four steps follows each warp's list (the kernel prefetches one batch ahead). ----
    const uint64_t half = (uint64_t)COLS * 8, keep = ~((uint64_t)COLS * 16 - 1);
    auto pack = [half, keep](double wv, int idx) {
        uint64_t b; std::memcpy(&b, &wv, 8);
        b = ((b + half) & keep) | ((uint64_t)idx << 4);
        double r; std::memcpy(&r, &b, 8); return r;
    };
    std::vector<double> ent;
    std::vector<int4> col_i((size_t)H * TC);
    std::vector<double2> col_d((size_t)H * TC * COLW);
    std::vector<double> eq0((size_t)H * TC, 1.0);
    std::vector<int32_t> col_ia((size_t)H * TC, -1);
    for (int wi = 0; wi < NW; ++wi)
        for (int j = 0; j <= H; ++j) {
            const size_t fbase = ent.size();
            if (j == H) { ent.resize(fbase + DCP_RES2_GB * 32, 0.0); break; }
            const int blk = slot_blk[wi][j], g = blk / npb, pb = blk % npb, K = ell_K[pb];
            ent.resize(fbase + (size_t)K * 32);
            for (int k = 0; k < K; ++k)
                for (int l = 0; l < 32; ++l) {
                    const size_t o = (size_t)ell_base[pb] + (size_t)k * 32 + l;
                    ent[fbase + (size_t)k * 32 + l] = pack(ell_w[o], g * nac_pad + ell_src[o]);
                }
            for (int l = 0; l < 32; ++l) {
                const int pos = pb * 32 + l, ia = thread_hru[pos], t = wi * 32 + l;
                const size_t o = (size_t)j * TC + t;
                const HruStatic *h = ia >= 0 ? &hru[ia] : nullptr;
                int4 ci;
                ci.x = (int)(fbase + l) | (K << 20);
                ci.y = (h ? h->ippt : 0) | ((h ? h->ipet : 0) << 8) | (g << 16) | (h ? 0 : (int)0x80000000u);
                ci.z = (g * nac_pad + pos) | ((g * nac_pad) << 16);
                ci.w = ((g * S * npt + (h ? h->ipar : 0)) * PARW) | ((g * nac_pad + rank[pos]) << 16);
                col_i[o] = ci;
                if (wide) {                                                                          // word-major: a warp's
                    col_d[o] = make_double2(h ? h->ac : 1.0, h ? cq[ia] : 0.0);                      // LDS.128 spans 512 B
                    col_d[(size_t)H * TC + o] = make_double2(h ? h->cosm : 1.0, h ? 1.0 / h->cosm : 1.0);
                    eq0[o] = h ? std::exp(-h->st) : 1.0;
                } else {
                    col_d[o] = make_double2(h ? h->ac : 1.0, h ? 1.0 / h->ac : 1.0);
                    col_d[(size_t)H * TC + o] = make_double2(h ? cq[ia] : 0.0, h ? h->cosm : 1.0);
                    col_d[(size_t)2 * H * TC + o] = make_double2(h ? 1.0 / h->cosm : 1.0, 0.0);
                }
                col_ia[o] = ia;
            }
        }
    if (getenv("DCP_RES_DEBUG"))
        fprintf(stderr, "[dcp plan2%s] gather list %zu entries (nnz %zu), K per block:%s; gather steps per warp:%s\n", wide ? "w" : "", ent.size(),
                w.size(), [&] { static std::string s2; s2.clear(); for (int pb = 0; pb < npb; ++pb) s2 += " " + std::to_string(ell_K[pb]); return s2.c_str(); }(),
                [&] { static std::string s3; s3.clear(); for (int wi = 0; wi < NW; ++wi) s3 += " " + std::to_string(load[wi]); return s3.c_str(); }());
    Q.ell_len = (int)ent.size();
    if (Q.ell_len >= (1 << 20)) return;

    Q.n_teams = G * nriv;
    Q.team = 32;
    while (Q.team > 1 && Q.team * Q.n_teams > TR) Q.team >>= 1;
    Q.n_pass = (Q.n_teams * Q.team + TR - 1) / TR;
    if (Q.n_pass > 16) return;
    // routing jobs of one member: gauge i x node u in {i} + upstream tree of i (order of route_process_run_step's sum,
    // DTA/dta_route_processing.f90:685-714)
    std::vector<int32_t> job_node, job_gauge, gauge_job_beg(nriv + 1, 0);
    {
        std::vector<int32_t> up_beg(d->nnode + 1, 0);
        for (int n = 0; n < d->nnode; ++n) up_beg[n + 1] = up_beg[n] + d->uptree_count[n];
        for (int i = 0; i < nriv; ++i) {                     // river / output index i == node i (flow_output_indexes)
            job_node.push_back(i); job_gauge.push_back(i);
            for (int e2 = up_beg[i]; e2 < up_beg[i + 1]; ++e2) { job_node.push_back(d->uptree_index[e2] - 1); job_gauge.push_back(i); }
            gauge_job_beg[i + 1] = (int)job_node.size();
        }
    }
    Q.jobs_per_member = (int)job_node.size();
    const int n_jobs = G * S * Q.jobs_per_member;
    if (n_jobs > 4096) return;
    // shared memory: [shared tables][team 0][team 1]
    size_t off = 0;
    auto take = [&](size_t bytes) { off = (off + 127) & ~(size_t)127; const size_t o = off; off += bytes; return (int)o; };
    Q.off_ellw = take((size_t)Q.ell_len * 8);
    Q.off_coli = take((size_t)H * TC * 16);
    Q.off_cold = take((size_t)H * TC * 16 * COLW);
    Q.off_rbeg = take((size_t)(nriv + 1) * 4);
    Q.off_eq0 = take(wide ? (size_t)H * TC * 8 : 0);
    Q.off_fmtab = take(wide ? 0 : (size_t)(FM_EXP_TAB + 2 * FM_LOG_TAB) * 8);
    Q.off_team = (int)((off + 127) & ~(size_t)127);
    off = 0;
    Q.off_qbf = take((size_t)2 * COLS * 16);
    Q.off_con = take((size_t)2 * COLS * 16);
    Q.off_cst = take((size_t)S * COLS * (wide ? 8 : 16));
    Q.off_par = take((size_t)G * S * npt * PARW * 16);
    Q.off_forc = take((size_t)2 * Q.Tt * (d->ngp + d->nge) * 8);
    Q.off_bar = take(16);
    Q.off_job = take((size_t)n_jobs * 16);
    Q.off_acc = take((size_t)G * S * nriv * 6 * 8);
    // narrow build: team blocks (tuple buffers first) sit on boundaries of one tuple buffer (the kernel ORs the tuple offset
    // into the buffer address) and the plan reserves the alignment slack; the wide build adds the offset instead (16 KB of
    // slack would not fit beside 1024 columns) and needs no alignment
    const size_t tupb = wide ? 128 : (size_t)COLS * 16;
    Q.team_bytes = (int)((off + tupb - 1) & ~(tupb - 1));
    Q.smem_bytes = Q.off_team + (wide ? 0 : (int)tupb) + (TEAMS - 1) * Q.team_bytes + (int)((off + 127) & ~(size_t)127);
    if ((size_t)Q.smem_bytes > c->smem_optin) return;
    std::vector<double> pet_pos(pet);
    for (double &x : pet_pos) x = x > 0.0 ? x : 0.0;
    auto up = [&](auto &vec, auto **dst) { if (e == cudaSuccess) e = upload(vec, dst, c->owned); };
    up(ent, &Q.ell_w); up(col_i, &Q.col_i); up(col_d, &Q.col_d); up(col_ia, &Q.col_ia);
    if (wide) up(eq0, &Q.eq0);
    {
        std::vector<double> fm_tab(FM_EXP_TAB + 2 * FM_LOG_TAB);
        fm_build_tables(fm_tab.data(), fm_tab.data() + FM_EXP_TAB);
        up(fm_tab, &Q.fm_tab);
    }
    up(rbeg, &Q.rbeg); up(pet_pos, &Q.pet_pos);
    up(job_node, &Q.job_node); up(job_gauge, &Q.job_gauge); up(gauge_job_beg, &Q.gauge_job_beg);
    std::vector<double> inv_frac(nriv), obs_eps(nriv), log_obs(qobs.size(), 0.0);
    for (int r = 0; r < nriv; ++r) { inv_frac[r] = 1.0 / frac[r]; obs_eps[r] = obs[r].eps; }
    for (size_t x = 0; x < qobs.size(); ++x) {
        const int g = (int)(x % nriv);
        if (qobs[x] != d->flow_err) log_obs[x] = std::log(qobs[x] > obs[g].eps ? qobs[x] : obs[g].eps);
    }
    up(inv_frac, &Q.inv_frac); up(obs_eps, &Q.obs_eps);
    if (!qobs.empty()) up(log_obs, &Q.log_obs); else Q.log_obs = nullptr;
    Q.feasible = 1;
}

}  // namespace

extern "C" {

int dcp_abi_version(void) { return DCP_ABI_VERSION; }

const char *dcp_last_error(void) { return g_err.c_str(); }

int dcp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int dcp_set_device(int device) {
    CUDA_TRY(cudaSetDevice(device));
    return DCP_OK;
}

int dcp_catchment_destroy(dcp_catchment *c) {
    if (!c) return DCP_OK;
    cudaSetDevice(c->device);
    for (void *p : c->owned) cudaFree(p);
    c->ws.release();
    c->io.release();
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return DCP_OK;
}

int dcp_catchment_create(const dcp_catchment_desc *d, dcp_catchment **out) {
    if (!d || !out) return fail(DCP_ERR_INVALID, "null argument");
    *out = nullptr;
    if (d->struct_size != (int)sizeof(dcp_catchment_desc))
        return fail(DCP_ERR_INVALID, "dcp_catchment_desc.struct_size %d != %d (ABI mismatch)", d->struct_size,
                    (int)sizeof(dcp_catchment_desc));
    if (dcp_device_count() < 1) return fail(DCP_ERR_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
    const int nac = d->nac, nriv = d->nriv, npt = d->npt, nnode = d->nnode, ncell = d->ncell;
    if (nac < 1 || nriv < 1 || npt < 1 || d->nstep < 1 || d->ngp < 1 || d->nge < 1 || d->ntt < 1 || !(d->dt > 0))
        return fail(DCP_ERR_INVALID, "non-positive dimension or dt");
    if (nnode < nriv || ncell < 1) return fail(DCP_ERR_INVALID, "routing: nnode (%d) < nriv (%d) or no river cells", nnode, nriv);
    if (d->qobs != nullptr && d->nstep_obs > 0 && (d->t_warm < 0 || d->t_warm > d->nstep))
        return fail(DCP_ERR_INVALID, "t_warm (%d) must lie in [0, nstep]", d->t_warm);
    if (!d->ac || !d->st || !d->mslp || !d->ippt || !d->ipet || !d->ipar || !d->ipriv || !d->ims_rz || !d->ims_uz ||
        !d->ntrans || !d->wtrans || !d->rivers || !d->sum_ac_riv || !d->qobs_start || !d->node_gauge_id ||
        !d->node_type || !d->uptree_count || !d->frac_sub_area || !d->river_data || !d->node_to_flow_mapping ||
        !d->rain || !d->pet)
        return fail(DCP_ERR_INVALID, "null array in descriptor");

    // ---------------- validate + flatten on the host ----------------
    std::vector<HruStatic> hru(nac);
    std::vector<std::vector<WEntry>> rows(nac);
    std::vector<RivEntry> riv;
    std::vector<uint8_t> layer_used(npt, 0);
    size_t tro = 0, wo = 0;
    for (int ia = 0; ia < nac; ++ia) {
        HruStatic &h = hru[ia];
        std::memset(&h, 0, sizeof(h));
        h.ac = d->ac[ia]; h.st = d->st[ia]; h.cosm = std::cos(d->mslp[ia]);
        h.ippt = d->ippt[ia] - 1; h.ipet = d->ipet[ia] - 1; h.ipar = d->ipar[ia] - 1; h.ipriv = d->ipriv[ia] - 1;
        h.ims_rz = d->ims_rz[ia]; h.ims_uz = d->ims_uz[ia];
        if (h.ippt < 0 || h.ippt >= d->ngp || h.ipet < 0 || h.ipet >= d->nge || h.ipar < 0 || h.ipar >= npt ||
            h.ipriv < 0 || h.ipriv >= nriv)
            return fail(DCP_ERR_INVALID, "HRU %d: ippt/ipet/ipar/ipriv out of range", ia + 1);
        layer_used[h.ipar] = 1;
        const int nt = d->ntrans[ia];
        if (nt < 0) return fail(DCP_ERR_INVALID, "HRU %d: ntrans < 0", ia + 1);
        if (nt > 0 && !d->itrans) return fail(DCP_ERR_INVALID, "itrans is null");
        for (int iw = 0; iw < nt; ++iw) {
            const int tgt = d->itrans[tro + iw] - 1;
            if (tgt < 0 || tgt >= nac) return fail(DCP_ERR_INVALID, "HRU %d: itrans(%d) out of range", ia + 1, iw + 1);
            WEntry e; e.w = d->wtrans[wo + iw]; e.ac_src = d->ac[ia]; e.src = ia; e.pad = 0;
            rows[tgt].push_back(e);        // sources arrive in ascending ia, iw: the reference's scatter order
        }
        h.w_riv = d->wtrans[wo + nt];
        tro += nt; wo += nt + 1;
        h.riv_beg = (int)riv.size();
        for (int r = 0; r < nriv; ++r) {
            const double sh = d->rivers[(size_t)r * nac + ia];
            if (sh != 0.0) { RivEntry e; e.share = sh; e.r = r; e.pad = 0; riv.push_back(e); }
        }
        h.riv_end = (int)riv.size();
    }
    std::vector<WEntry> w;
    for (int ia = 0; ia < nac; ++ia) {
        hru[ia].row_beg = (int)w.size();
        w.insert(w.end(), rows[ia].begin(), rows[ia].end());
        hru[ia].row_end = (int)w.size();
    }

    // river network (DTA/dta_route_processing.f90:184-243, member-independent parts)
    int n_gauge = 0;
    for (int n = 0; n < nnode; ++n) if (d->node_type[n] == 2) ++n_gauge;
    if (n_gauge != nriv)   // size check of route_process_run_step (:663-668) would STOP
        return fail(DCP_ERR_INVALID, "routing: %d gauge nodes but %d rivers (reference stops: size mismatch)", n_gauge, nriv);
    auto rd = [&](int cell, int col) { return d->river_data[(size_t)(col - 1) * ncell + cell]; };
    double min_outlet = rd(0, 7);
    for (int j = 1; j < ncell; ++j) min_outlet = std::min(min_outlet, rd(j, 7));
    std::vector<int32_t> cbeg(nnode, -1), cend(nnode, -1);
    std::vector<double> reach_dist(nnode, 0.0), down_dist(nnode, 0.0), cell_a(ncell), cell_b(ncell), cell_area(ncell);
    for (int n = 0; n < nnode; ++n) {
        int s = 0, e = ncell;
        for (int j = 1; j <= ncell; ++j) {
            const int id = (int)std::lround(rd(j - 1, 1));
            if (s == 0) { if (id == d->node_gauge_id[n]) s = j; }
            else if (id != d->node_gauge_id[n]) { e = j - 1; break; }
        }
        if (s == 0)
            return fail(DCP_ERR_UNSUPPORTED, "routing node %d (id %d) has no river cells (stale delays upstream, "
                        "dta_route_processing.f90:204-212)", n + 1, d->node_gauge_id[n]);
        cbeg[n] = s - 1; cend[n] = e;
        double mx = rd(s - 1, 3), mn = rd(s - 1, 7);
        for (int j = s; j <= e; ++j) { mx = std::max(mx, rd(j - 1, 3)); mn = std::min(mn, rd(j - 1, 7)); }
        reach_dist[n] = mx - mn;
        down_dist[n] = mn - min_outlet;
    }
    for (int j = 0; j < ncell; ++j) {
        const double cell_dist = rd(j, 3) - rd(j, 7);          // :325
        cell_a[j] = rd(j, 4);                                  // :333
        cell_b[j] = cell_a[j] + cell_dist;                     // :334
        cell_area[j] = rd(j, 2);
        if (cell_a[j] < 0 || cell_b[j] < 0) return fail(DCP_ERR_INVALID, "river cell %d: negative section distance", j + 1);
    }
    std::vector<int32_t> node_input(nnode, -1), up_beg(nnode + 1, 0), up_idx;
    std::vector<double> frac(nriv);
    for (int r = 0; r < nriv; ++r) {
        const int n = d->node_to_flow_mapping[r] - 1;
        if (n < 0 || n >= nnode) return fail(DCP_ERR_INVALID, "node_to_flow_mapping(%d) out of range", r + 1);
        node_input[n] = r;
        frac[r] = d->frac_sub_area[n];
    }
    for (int n = 0; n < nnode; ++n) {
        const int cnt = d->uptree_count[n];
        if (cnt < 0 || (cnt > 0 && !d->uptree_index)) return fail(DCP_ERR_INVALID, "bad upstream tree");
        for (int k = 0; k < cnt; ++k) {
            const int u = d->uptree_index[up_beg[n] + k] - 1;
            if (u < 0 || u >= nnode) return fail(DCP_ERR_INVALID, "upstream index out of range");
            up_idx.push_back(u);
        }
        up_beg[n + 1] = up_beg[n] + cnt;
    }

    // member-independent reach totals (dyna_write_output.f90:61-77; sum_pptin/sum_pein of dyna_root_zone1.f90:40,63)
    std::vector<double> sum_p(d->ngp, 0.0), sum_e(d->nge, 0.0), tot_ppt(nriv, 0.0), tot_pet(nriv, 0.0), tot_frac(nriv, 0.0);
    for (int g = 0; g < d->ngp; ++g)
        for (int t = 0; t < d->nstep; ++t) sum_p[g] = sum_p[g] + d->rain[(size_t)t * d->ngp + g];
    for (int g = 0; g < d->nge; ++g)
        for (int t = 0; t < d->nstep; ++t) {
            const double ep = d->pet[(size_t)t * d->nge + g];
            if (ep > 0) sum_e[g] = sum_e[g] + ep;
        }
    for (int ia = 0; ia < nac; ++ia) {
        const HruStatic &h = hru[ia];
        const double sp = h.ims_rz == 1 ? sum_p[h.ippt] : 0.0, se = h.ims_rz == 1 ? sum_e[h.ipet] : 0.0;
        tot_ppt[h.ipriv] = tot_ppt[h.ipriv] + (sp * h.ac);
        tot_pet[h.ipriv] = tot_pet[h.ipriv] + (se * h.ac);
        tot_frac[h.ipriv] = tot_frac[h.ipriv] + h.ac;
    }

    // observed-flow statistics of the NSE / log-NSE extension (definition: DESIGN.md "Objective")
    std::vector<ObsStat> obs(nriv);
    std::vector<double> qobs;
    const bool has_qobs = d->qobs != nullptr && d->nstep_obs > 0;
    if (has_qobs) {
        qobs.assign(d->qobs, d->qobs + (size_t)nriv * d->nstep_obs);
        const int n_t = std::min(d->nstep, d->nstep_obs);
        auto lclip = [](double x, double eps) { return std::log(x > eps ? x : eps); };
        for (int g = 0; g < nriv; ++g) {
            ObsStat s{};
            double sum = 0;
            for (int t = d->t_warm; t < n_t; ++t) {
                const double o = qobs[(size_t)t * nriv + g];
                if (o != d->flow_err) { sum = sum + o; s.n++; }
            }
            s.sum = sum;
            if (s.n > 0) {
                s.mean = sum / s.n; s.eps = s.mean / 100;
                double lsum = 0;
                for (int t = d->t_warm; t < n_t; ++t) {
                    const double o = qobs[(size_t)t * nriv + g];
                    if (o != d->flow_err) { s.sst = s.sst + (o - s.mean) * (o - s.mean); lsum = lsum + lclip(o, s.eps); }
                }
                s.lmean = lsum / s.n;
                for (int t = d->t_warm; t < n_t; ++t) {
                    const double o = qobs[(size_t)t * nriv + g];
                    if (o != d->flow_err) { const double lo = lclip(o, s.eps); s.lsst = s.lsst + (lo - s.lmean) * (lo - s.lmean); }
                }
            }
            obs[g] = s;
        }
    }

    // ---------------- upload ----------------
    dcp_catchment *c = new dcp_catchment();
    cudaGetDevice(&c->device);
    DevCatchment &C = c->C;
    C.nac = nac; C.nriv = nriv; C.npt = npt; C.nstep = d->nstep; C.ngp = d->ngp; C.nge = d->nge; C.ntt = d->ntt;
    C.nnode = nnode; C.ncell = ncell;
    C.dt = d->dt; C.dtt = d->dt / (double)d->ntt; C.log_dt = std::log(d->dt); C.inv_dtt = 1.0 / C.dtt;
    C.nstep_obs = has_qobs ? d->nstep_obs : 0; C.t_warm = d->t_warm; C.flow_err = d->flow_err;
    c->has_qobs = has_qobs;
    c->node_reach_dist = reach_dist; c->node_down_dist = down_dist;
    // forcing zero-padded to whole TMA tiles (the resident kernel bulk-copies DCP_RES_TT steps at a time)
    const size_t nstep_pad = ((size_t)d->nstep + DCP_FORCING_PAD - 1) / DCP_FORCING_PAD * DCP_FORCING_PAD + DCP_FORCING_PAD;
    std::vector<double> rain((size_t)d->ngp * nstep_pad, 0.0), pet((size_t)d->nge * nstep_pad, 0.0);
    std::copy(d->rain, d->rain + (size_t)d->ngp * d->nstep, rain.begin());
    std::copy(d->pet, d->pet + (size_t)d->nge * d->nstep, pet.begin());
    std::vector<double> sum_ac_riv(d->sum_ac_riv, d->sum_ac_riv + nriv), qobs_start(d->qobs_start, d->qobs_start + nriv);
    cudaError_t e = cudaSuccess;
    auto up = [&](auto &vec, auto **dst) { if (e == cudaSuccess) e = upload(vec, dst, c->owned); };
    up(hru, &C.hru); up(w, &C.w); up(riv, &C.riv); up(rain, &C.rain); up(pet, &C.pet);
    up(sum_ac_riv, &C.sum_ac_riv); up(qobs_start, &C.qobs_start); up(layer_used, &C.layer_used);
    up(cbeg, &C.node_cell_beg); up(cend, &C.node_cell_end); up(reach_dist, &C.node_reach_dist);
    up(down_dist, &C.node_down_dist); up(cell_a, &C.cell_a); up(cell_b, &C.cell_b); up(cell_area, &C.cell_area);
    up(node_input, &C.node_input); up(up_beg, &C.up_beg); up(up_idx, &C.up_idx); up(frac, &C.frac_sub_area);
    { std::vector<double> ones(nriv, 1.0); up(ones, &c->ones_riv); }
    up(obs, &C.obs); up(tot_ppt, &C.tot_ppt); up(tot_pet, &C.tot_pet); up(tot_frac, &C.tot_frac);
    if (has_qobs) up(qobs, &C.qobs); else C.qobs = nullptr;
    // ---------------- resident-kernel plan ----------------
    {
        cudaDeviceProp prop{};
        cudaGetDeviceProperties(&prop, c->device);
        c->n_sm = prop.multiProcessorCount;
        c->smem_optin = prop.sharedMemPerBlockOptin;
        ResPlan &P = c->plan;
        P = ResPlan{};
        int nac_pad = 32;
        while (nac_pad < nac) nac_pad <<= 1;
        const int S = DCP_RES_S, CT = DCP_RES_COMPUTE;
        if (nac_pad > CT) {
            c->plan_why = "nac > 512 HRUs (513 .. 1024: FAST policy only, wide resident kernel)";
            // ---- wide build of the second-generation kernel: one team of 1024 columns (dcp_resident_v2w.cu) ----
            c->plan2w = Res2Plan{};
            bool lean_ok = d->ntt == 1;
            for (int i = 0; i < nac; ++i) if (hru[i].ims_rz != 1 || hru[i].ims_uz != 1) lean_ok = false;
            if (nac <= 1024 && lean_ok && !getenv("DCP_RES_NO_V2")) {
                std::vector<int32_t> order(nac), thread_hru(1024, -1), hru_pos(nac, 0);
                for (int i = 0; i < nac; ++i) order[i] = i;
                std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
                    return (hru[a].row_end - hru[a].row_beg) > (hru[b].row_end - hru[b].row_beg); });
                for (int i = 0; i < nac; ++i) { thread_hru[i] = order[i]; hru_pos[order[i]] = i; }
                build_plan2(c, d, hru, w, thread_hru, hru_pos, 1024, pet, frac, qobs, obs, true, e);
            }
        } else {
            P.nac_pad = nac_pad; P.G = CT / nac_pad; P.Tt = DCP_RES_TT; P.nnz = (int)w.size();
            while (P.G > 1 && P.G * S * nriv > DCP_RES_MAXPAIRS) P.G >>= 1;    // river lanes carry 2 chains per (member, river)
            P.Mc = P.G * S;
            std::vector<RivChain> rivq;
            std::vector<int32_t> rivq_beg(nriv + 1, 0), rivo, rivo_beg(nriv + 1, 0);
            for (int r = 0; r < nriv; ++r) {
                for (int ia = 0; ia < nac; ++ia) {
                    const double sh = d->rivers[(size_t)r * nac + ia];
                    if (sh != 0.0) { RivChain rc{}; rc.w_riv = hru[ia].w_riv; rc.share = sh; rc.ac = hru[ia].ac; rc.ia = ia; rivq.push_back(rc); }
                    if (hru[ia].ipriv == r) rivo.push_back(ia);
                }
                rivq_beg[r + 1] = (int)rivq.size();
                rivo_beg[r + 1] = (int)rivo.size();
            }
            P.nq = (int)rivq.size();
            const int n_chains = 2 * P.Mc * nriv;
            P.n_threads = CT + 128;
            P.variant = getenv("DCP_RES_VARIANT") ? atoi(getenv("DCP_RES_VARIANT")) : 0;   // 0 = chosen per arithmetic policy
            P.river_team = 1;
            while (P.river_team < 16 && 2 * P.river_team * n_chains <= 96) P.river_team <<= 1;      // 96 summing lanes (3 warps)
            size_t off = 0;
            auto take = [&](size_t bytes) { off = (off + 127) & ~(size_t)127; const size_t o = off; off += bytes; return (int)o; };
            P.off_qbf = take((size_t)2 * P.G * nac_pad * S * 8);
            P.off_ex = take((size_t)2 * S * CT * 8);
            P.off_cst = take((size_t)5 * S * CT * 8);
            P.off_par = take((size_t)P.Mc * npt * 6 * 8);
            P.off_w = take((size_t)P.nnz * 8);
            P.off_src = take((size_t)P.nnz * 2);
            P.off_ac = take((size_t)nac_pad * 8);
            P.off_forc = take((size_t)2 * P.Tt * (d->ngp + d->nge) * 8);
            P.off_bar = take(16);
            P.off_rqc = take((size_t)3 * (P.nq + 1) * 8);
            P.off_rq = take((size_t)P.nq * 2);
            P.off_ro = take((size_t)nac * 2);
            P.smem_bytes = (int)((off + 127) & ~(size_t)127);
            if (P.Mc * nriv > DCP_RES_MAXPAIRS) c->plan_why = "too many rivers for the river warps";
            else if ((size_t)P.smem_bytes > c->smem_optin) {
                char buf[160];
                snprintf(buf, sizeof(buf), "needs %d B of shared memory per CTA (W has %d entries), device allows %zu",
                         P.smem_bytes, P.nnz, c->smem_optin);
                c->plan_why = buf;
            } else {
                // thread positions sorted by W row length (descending): the lanes of a warp walk rows of similar length
                std::vector<int32_t> order(nac), thread_hru(nac_pad, -1), hru_pos(nac, 0);
                for (int i = 0; i < nac; ++i) order[i] = i;
                std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
                    return (hru[a].row_end - hru[a].row_beg) > (hru[b].row_end - hru[b].row_beg); });
                for (int i = 0; i < nac; ++i) { thread_hru[i] = order[i]; hru_pos[order[i]] = i; }
                up(thread_hru, &P.thread_hru); up(hru_pos, &P.hru_pos); up(rivq, &P.rivq); up(rivq_beg, &P.rivq_beg);
                up(rivo, &P.rivo); up(rivo_beg, &P.rivo_beg);
                P.feasible = 1;
                P.lean_ok = d->ntt == 1;
                for (int i = 0; i < nac; ++i) if (hru[i].ims_rz != 1 || hru[i].ims_uz != 1) P.lean_ok = 0;

                // ---- second-generation plan (dcp_resident_v2.cu) ----
                c->plan2 = Res2Plan{};
                if (P.lean_ok && !getenv("DCP_RES_NO_V2"))
                    build_plan2(c, d, hru, w, thread_hru, hru_pos, nac_pad, pet, frac, qobs, obs, false, e);
            }
        }
    }
    // ---------------- stepwise plan: ordered river lists, dense form of W when it is dense enough ----------------
    {
        DensePlan &D = c->dplan;
        D = DensePlan{};
        std::vector<int32_t> rq_beg(nriv + 1, 0), rq_idx, ro_beg(nriv + 1, 0), ro_idx;
        for (int r = 0; r < nriv; ++r) {
            for (int ia = 0; ia < nac; ++ia) {
                for (int x = hru[ia].riv_beg; x < hru[ia].riv_end; ++x) if (riv[x].r == r) rq_idx.push_back(x);
                if (hru[ia].ipriv == r) ro_idx.push_back(ia);
            }
            rq_beg[r + 1] = (int)rq_idx.size(); ro_beg[r + 1] = (int)ro_idx.size();
        }
        if (rq_idx.empty()) rq_idx.push_back(0);
        if (ro_idx.empty()) ro_idx.push_back(0);
        D.nq = (int)riv.size();
        up(rq_beg, &D.rq_beg); up(rq_idx, &D.rq_idx); up(ro_beg, &D.ro_beg); up(ro_idx, &D.ro_idx);
        const double density = (double)w.size() / ((double)nac * nac);
        const double min_density = getenv("DCP_DENSE_MIN_DENSITY") ? atof(getenv("DCP_DENSE_MIN_DENSITY")) : 0.05;
        D.Mp = (nac + 127) / 128 * 128; D.Kp = (nac + 15) / 16 * 16;
        const size_t dense_bytes = (size_t)D.Mp * D.Kp * 8;
        if (density >= min_density && nac >= 128 && dense_bytes <= ((size_t)2 << 30)) {
            std::vector<double> wd((size_t)D.Mp * D.Kp, 0.0), acs(D.Kp, 0.0);
            bool dup = false;
            for (int ia = 0; ia < nac; ++ia) {
                acs[ia] = hru[ia].ac;
                for (int x = hru[ia].row_beg; x < hru[ia].row_end; ++x) {
                    double &cell = wd[(size_t)ia * D.Kp + w[x].src];
                    if (cell != 0.0) dup = true;                  // two links between the same pair: keep the sparse form
                    cell = w[x].w;
                }
            }
            std::vector<int32_t> kt_beg(D.Mp / 128 + 1, 0), kt_idx;
            for (int bm = 0; bm < D.Mp / 128; ++bm) {
                for (int kt = 0; kt < D.Kp / 16; ++kt) {
                    bool any = false;
                    for (int r = bm * 128; r < bm * 128 + 128 && !any; ++r)
                        for (int k = kt * 16; k < kt * 16 + 16; ++k) if (wd[(size_t)r * D.Kp + k] != 0.0) { any = true; break; }
                    if (any) kt_idx.push_back(kt);
                }
                kt_beg[bm + 1] = (int)kt_idx.size();
            }
            if (kt_idx.empty()) kt_idx.push_back(0);
            // destination blocks longest-first, so that the blocks with few source tiles fill the tail of the grid
            std::vector<int32_t> bm_order(D.Mp / 128);
            for (int i = 0; i < D.Mp / 128; ++i) bm_order[i] = i;
            std::stable_sort(bm_order.begin(), bm_order.end(), [&](int a, int b) {
                return kt_beg[a + 1] - kt_beg[a] > kt_beg[b + 1] - kt_beg[b]; });
            if (!dup) {
                up(bm_order, &D.bm_order);
                up(wd, &D.wd); up(acs, &D.acs); up(kt_beg, &D.kt_beg); up(kt_idx, &D.kt_idx);
                D.dense = 1;
            }
        }
        D.feasible = 1;
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        const int code = fail(e == cudaErrorMemoryAllocation ? DCP_ERR_NOMEM : DCP_ERR_CUDA, "catchment upload failed: %s",
                              cudaGetErrorString(e));
        dcp_catchment_destroy(c);
        return code;
    }
    *out = c;
    return DCP_OK;
}

int dcp_get_run_stats(const dcp_catchment *c, dcp_run_stats *out) {
    if (!c || !out) return fail(DCP_ERR_INVALID, "null argument");
    *out = c->stats;
    return DCP_OK;
}

}  // extern "C"

namespace {

struct Carver {                     // carves a slab into aligned arrays
    char *base; size_t off = 0;
    explicit Carver(void *p) : base((char *)p) {}
    template <class T> T *take(size_t n) {
        off = (off + 255) & ~(size_t)255;
        T *r = base ? (T *)(base + off) : nullptr;
        off += n * sizeof(T);
        return r;
    }
};

// lays the batch workspace out (base == nullptr: size query only)
// ring_slots > 0: fused routing (dcp_route_fused.cuh) -- qout / y are rings of R steps per member IN FLIGHT, and the
// per-step objective series do not exist
size_t carve_batch(void *base, const DevCatchment &C, DevBatch &W, bool check, bool aet, int stepwise_nq = -1, int ring_slots = 0,
                   bool want_sig = false) {
    Carver cv(base);
    const size_t B = W.B, nac = C.nac;
    W.par = cv.take<double>(7 * C.npt * B);
    W.t0dt = cv.take<double>(C.npt * B);
    W.sd = cv.take<double>(nac * B); W.suz = cv.take<double>(nac * B); W.srz = cv.take<double>(nac * B);
    W.qint1 = cv.take<double>(nac * B); W.qbf0 = cv.take<double>(nac * B); W.qbf1 = cv.take<double>(nac * B);
    W.q0 = cv.take<double>(nac * B); W.q2 = cv.take<double>(nac * B);
    W.aet = (aet || check) ? cv.take<double>(nac * B) : nullptr;
    if (check) {
        W.sum_pptin = cv.take<double>(nac * B); W.sum_pein = cv.take<double>(nac * B);
        W.sumexs = cv.take<double>(nac * B); W.sumexus = cv.take<double>(nac * B);
        W.sdini = cv.take<double>(nac * B); W.srzini = cv.take<double>(nac * B); W.qbfini = cv.take<double>(nac * B);
        W.riv_sums = cv.take<double>(3 * C.nriv * B);
    } else {
        W.sum_pptin = W.sum_pein = W.sumexs = W.sumexus = W.sdini = W.srzini = W.qbfini = W.riv_sums = nullptr;
    }
    W.dh = cv.take<double>((size_t)C.nnode * W.Lcap * B);
    W.hstart = cv.take<int32_t>(C.nnode * B);
    W.hbins = cv.take<int32_t>(C.nnode * B);
    W.qout = cv.take<double>((size_t)W.R * C.nriv * (ring_slots > 0 ? (size_t)ring_slots : B));
    W.y = cv.take<double>((size_t)W.R * C.nnode * (ring_slots > 0 ? (size_t)ring_slots : B));
    W.q_user = nullptr;
    if (C.qobs && ring_slots <= 0) {
        W.d2 = cv.take<double>((size_t)W.R * C.nriv * B); W.dl2 = cv.take<double>((size_t)W.R * C.nriv * B);
        W.qs = cv.take<double>((size_t)W.R * C.nriv * B);
    } else { W.d2 = W.dl2 = W.qs = nullptr; }
    W.sse = cv.take<double>(C.nriv * B); W.lsse = cv.take<double>(C.nriv * B);
    W.sq = cv.take<double>(C.nriv * B); W.sqq = cv.take<double>(C.nriv * B); W.sqo = cv.take<double>(C.nriv * B);
    W.bad = cv.take<int32_t>(C.nriv * B);
    W.fdc = want_sig ? cv.take<uint32_t>((size_t)DCP_FDC_BINS * C.nriv * B) : nullptr;
    W.status = cv.take<int32_t>(B);
    W.diag = cv.take<double>(3 * B);
    if (stepwise_nq >= 0) {
        W.qin_ext = cv.take<double>(nac * B); W.qof_term = cv.take<double>(nac * B);
        W.qb_term = cv.take<double>((size_t)std::max(stepwise_nq, 1) * B);
        W.bad_src = cv.take<int32_t>(B);
        W.t_dev = cv.take<int32_t>(64);
    } else { W.qin_ext = W.qof_term = W.qb_term = nullptr; W.bad_src = nullptr; W.t_dev = nullptr; }
    return cv.off + 256;
}

struct EventTimer {                 // sums the elapsed time of many (start, stop) pairs on one stream
    std::vector<cudaEvent_t> ev;
    std::vector<int> tag;
    cudaStream_t s;
    explicit EventTimer(cudaStream_t st) : s(st) {}
    void mark(int t) {
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s);
        ev.push_back(e); tag.push_back(t);
    }
    // consecutive marks delimit intervals; interval i carries tag[i]
    void collect(double *sums, int ntags) {
        for (int i = 0; i < ntags; ++i) sums[i] = 0;
        for (size_t i = 0; i + 1 < ev.size(); ++i) {
            float ms = 0; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            if (tag[i] >= 0 && tag[i] < ntags) sums[tag[i]] += ms;
        }
    }
    double total() { float ms = 0; if (ev.size() >= 2) cudaEventElapsedTime(&ms, ev.front(), ev.back()); return ms; }
    ~EventTimer() { for (auto e : ev) cudaEventDestroy(e); }
};
enum { T_H2D = 0, T_INIT, T_PHYS, T_ROUTE, T_FIN, T_D2H, T_OTHER, T_NTAGS };

}  // namespace

extern "C" {

int dcp_run_ensemble(dcp_catchment *c, int64_t first_member, int64_t n_member, double *mcpar,
                     const dcp_run_opts *opts, double *q_out, double *perf_out, double *reach_totals,
                     double *init_diag, int32_t *status) {
    (void)first_member;
    if (!c || !mcpar) return fail(DCP_ERR_INVALID, "null handle or mcpar");
    if (n_member < 1) return fail(DCP_ERR_INVALID, "n_member < 1");
    dcp_run_opts o{};
    o.struct_size = sizeof(dcp_run_opts);
    if (opts) {
        if (opts->struct_size != (int)sizeof(dcp_run_opts)) return fail(DCP_ERR_INVALID, "dcp_run_opts.struct_size mismatch");
        o = *opts;
    }
    CUDA_TRY(cudaSetDevice(c->device));
    const DevCatchment &C = c->C;
    const bool dev_io = (o.flags & DCP_OPT_DEVICE_BUFFERS) != 0;
    const bool check = (o.flags & DCP_OPT_CHECK_BALANCE) != 0;
    const bool fast = (o.flags & DCP_OPT_FAST_MATH) != 0;
    const bool aet = reach_totals != nullptr;
    const bool want_perf = perf_out != nullptr;
    const bool want_sig = (o.flags & DCP_OPT_SIGNATURES) != 0;
    if (want_sig && !o.sig_out) return fail(DCP_ERR_INVALID, "DCP_OPT_SIGNATURES without dcp_run_opts.sig_out");
    // kernel choice: the resident kernel whenever the catchment fits on chip; the balance checks
    // (results_ts / root_zone1 STOP conditions) exist in the streamed kernel only
    bool use_res = false, use_sw = false;
    const bool res_ok = c->plan.feasible || (fast && c->plan2w.feasible);     // 513 .. 1024 HRUs: the wide kernel, FAST policy
    // The balance STOP conditions (results_ts, root_zone1) are statements about the reference's own rounding: they are
    // evaluated by the streamed kernel in the EXACT policy.  AUTO runs checked ensembles there; an explicit resident /
    // stepwise request keeps its kernel for every output and adds ONE streamed pass that only contributes the flag bits.
    if (check && (o.kernel == DCP_KERNEL_RESIDENT || o.kernel == DCP_KERNEL_STEPWISE)) {
        if (!status) return fail(DCP_ERR_INVALID, "DCP_OPT_CHECK_BALANCE without a status array");
        const size_t nb = (size_t)C.npt * 7 * (size_t)n_member * 8;
        dcp_run_opts om = o, oc = o;
        om.flags &= ~DCP_OPT_CHECK_BALANCE;
        oc.kernel = DCP_KERNEL_STREAMED;
        oc.flags = (o.flags & DCP_OPT_DEVICE_BUFFERS) | DCP_OPT_CHECK_BALANCE;
        oc.sig_out = nullptr;
        const int32_t bits = DCP_ST_RZ_BALANCE | DCP_ST_TS_OUTSUM | DCP_ST_TS_WBAL;
        int rc;
        if (dev_io) {                           // the main pass clips srinit in place: the flag pass gets the caller's values
            DevBuf tmp, st2;
            CUDA_TRY(tmp.reserve(nb));
            CUDA_TRY(st2.reserve((size_t)n_member * 4));
            CUDA_TRY(cudaMemcpy(tmp.p, mcpar, nb, cudaMemcpyDeviceToDevice));
            rc = dcp_run_ensemble(c, first_member, n_member, mcpar, &om, q_out, perf_out, reach_totals, init_diag, status);
            if (rc != DCP_OK) return rc;
            const dcp_run_stats main_stats = c->stats;
            rc = dcp_run_ensemble(c, first_member, n_member, (double *)tmp.p, &oc, nullptr, nullptr, nullptr, nullptr, (int32_t *)st2.p);
            if (rc != DCP_OK) return rc;
            dcp_launch_or_status(status, (const int32_t *)st2.p, bits, n_member, c->stream);
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            const double ms_chk = c->stats.ms_total;
            c->stats = main_stats; c->stats.ms_total += ms_chk; c->stats.n_launches += 1;
        } else {
            std::vector<double> tmp(mcpar, mcpar + nb / 8);
            std::vector<int32_t> st2((size_t)n_member);
            rc = dcp_run_ensemble(c, first_member, n_member, mcpar, &om, q_out, perf_out, reach_totals, init_diag, status);
            if (rc != DCP_OK) return rc;
            const dcp_run_stats main_stats = c->stats;
            rc = dcp_run_ensemble(c, first_member, n_member, tmp.data(), &oc, nullptr, nullptr, nullptr, nullptr, st2.data());
            if (rc != DCP_OK) return rc;
            for (int64_t m = 0; m < n_member; ++m) status[m] |= st2[m] & bits;
            const double ms_chk = c->stats.ms_total;
            c->stats = main_stats; c->stats.ms_total += ms_chk;
        }
        return DCP_OK;
    }
    if (o.kernel == DCP_KERNEL_RESIDENT) {
        if (!res_ok) return fail(DCP_ERR_UNSUPPORTED, "resident kernel unavailable: %s", c->plan_why.c_str());
        use_res = true;
    } else if (o.kernel == DCP_KERNEL_AUTO) {
        use_res = res_ok && !check;
        use_sw = !use_res && !check && c->dplan.feasible;      // catchments too large for the chip: one launch set per step
    } else if (o.kernel == DCP_KERNEL_STEPWISE) {
        if (!c->dplan.feasible) return fail(DCP_ERR_UNSUPPORTED, "stepwise kernels unavailable for this catchment");
        use_sw = true;
    } else if (o.kernel != DCP_KERNEL_STREAMED) {
        return fail(DCP_ERR_INVALID, "unknown kernel selector %d", o.kernel);
    }
    cudaStream_t s = c->stream;
    const size_t par_per_member = (size_t)C.npt * 7;

    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    free_b += c->ws.cap + c->io.cap;            // our own slabs are reusable
    const double mem_budget = 0.85 * (double)free_b;
    // time tile: the resident kernel runs a member group through ALL steps in one launch
    int tile = use_res ? C.nstep : (o.steps_per_tile > 0 ? o.steps_per_tile : 1024);
    tile = std::min(tile, C.nstep);
    const bool use_v2w = use_res && !c->plan.feasible;          // wide build: 2 members per CTA
    const int wave = use_res ? c->n_sm * (use_v2w ? 2 : c->plan.Mc) : 32;       // members that fill the GPU once
    int64_t Bmax = o.members_per_batch > 0 ? o.members_per_batch : (use_res ? (int64_t)wave * 64 : 131072);
    Bmax = std::min<int64_t>(Bmax, n_member);

    // resident kernels route and score inside the kernel (dcp_route_fused.cuh): rings per member in flight instead of
    // nstep-long series per member.  (The two-column lean variant of the first generation keeps the separate kernels.)
    const bool use_v2 = use_res && fast && (use_v2w || c->plan2.feasible);
    const bool v1_two_col = use_res && !use_v2 && fast && c->plan.lean_ok && (c->plan.variant ? c->plan.variant : 2) == 2;
    const bool fused = use_res && !v1_two_col;
    const int ring_slots = !fused ? 0 : use_v2w ? c->n_sm * 2 : use_v2 ? c->n_sm * 2 * c->plan2.G * 2 : c->n_sm * c->plan.Mc;

    c->stats = dcp_run_stats{};
    c->stats.kernel_used = use_res ? DCP_KERNEL_RESIDENT : use_sw ? DCP_KERNEL_STEPWISE : DCP_KERNEL_STREAMED;
    const int sw_nq = use_sw ? c->dplan.nq : -1;
    const bool sw_dense = use_sw && c->dplan.dense && !getenv("DCP_STEPWISE_NO_GEMM");
    EventTimer tm(s);
    int64_t n_launch = 0, n_phys = 0;
    int hist_len_max = 0;
    DevBuf vmin_buf;
    CUDA_TRY(vmin_buf.reserve(sizeof(double)));

    auto io_size = [&](int64_t B) {
        size_t b = 256;
        if (!dev_io) {
            b += par_per_member * B * 8 + 256;
            if (q_out) b += (size_t)C.nriv * C.nstep * B * 8 + 256;
            if (want_perf) b += (size_t)DCP_NMEASURE * C.nriv * B * 8 + 256;
            if (want_sig) b += (size_t)DCP_NSIG * C.nriv * B * 8 + 256;
            if (reach_totals) b += (size_t)3 * C.nriv * B * 8 + 256;
            if (init_diag) b += (size_t)3 * B * 8 + 256;
            if (status) b += (size_t)B * 4 + 256;
        }
        return b;
    };
    auto shrink = [&](int64_t B) {              // next smaller batch, whole waves when possible
        int64_t nb = B / 2;
        if (nb > wave) nb = nb / wave * wave;
        return std::max<int64_t>(1, nb);
    };

    for (int64_t m0 = 0; m0 < n_member;) {
        int64_t Bn = std::min<int64_t>(Bmax, n_member - m0);
        // ---- first guess of the batch size: everything except the (still unknown) histogram length ----
        for (;;) {
            DevBatch Wq{}; Wq.B = (int)Bn; Wq.Lcap = 1; Wq.R = fused ? 512 : use_res ? C.nstep : tile + 64;
            const size_t fixed = carve_batch(nullptr, C, Wq, check, aet, sw_nq, ring_slots, want_sig) + io_size(Bn);
            if ((double)fixed < 0.9 * mem_budget || Bn <= 1) break;
            Bn = shrink(Bn);
        }
        const size_t io_bytes = io_size(Bn);
        CUDA_TRY(c->io.reserve(io_bytes));
        tm.mark(T_H2D);
        Carver iocv(c->io.p);
        double *d_mcpar = nullptr, *d_q = nullptr, *d_perf = nullptr, *d_tot = nullptr, *d_diag = nullptr, *d_sig = nullptr;
        int32_t *d_status = nullptr;
        if (dev_io) {
            d_mcpar = mcpar + m0 * par_per_member;
            d_q = q_out ? q_out + (size_t)m0 * C.nriv * C.nstep : nullptr;
            d_perf = perf_out ? perf_out + (size_t)m0 * DCP_NMEASURE * C.nriv : nullptr;
            d_sig = want_sig ? o.sig_out + (size_t)m0 * DCP_NSIG * C.nriv : nullptr;
            d_tot = reach_totals ? reach_totals + (size_t)m0 * 3 * C.nriv : nullptr;
            d_diag = init_diag ? init_diag + (size_t)m0 * 3 : nullptr;
            d_status = status ? status + m0 : nullptr;
        } else {
            d_mcpar = iocv.take<double>(par_per_member * Bn);
            if (q_out) d_q = iocv.take<double>((size_t)C.nriv * C.nstep * Bn);
            if (want_perf) d_perf = iocv.take<double>((size_t)DCP_NMEASURE * C.nriv * Bn);
            if (want_sig) d_sig = iocv.take<double>((size_t)DCP_NSIG * C.nriv * Bn);
            if (reach_totals) d_tot = iocv.take<double>((size_t)3 * C.nriv * Bn);
            if (init_diag) d_diag = iocv.take<double>((size_t)3 * Bn);
            if (status) d_status = iocv.take<int32_t>(Bn);
            CUDA_TRY(cudaMemcpyAsync(d_mcpar, mcpar + m0 * par_per_member, par_per_member * Bn * 8,
                                     cudaMemcpyHostToDevice, s));
        }
        // ---- slowest channel velocity of the batch sizes the histograms ----
        tm.mark(T_OTHER);
        const unsigned long long big = 0x7fefffffffffffffull;
        CUDA_TRY(cudaMemcpyAsync(vmin_buf.p, &big, 8, cudaMemcpyHostToDevice, s));
        dcp_launch_min_chv(d_mcpar, C.npt, Bn, (double *)vmin_buf.p, s); ++n_launch;
        double chv_min = 0;
        CUDA_TRY(cudaMemcpyAsync(&chv_min, vmin_buf.p, 8, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        const double v_min = chv_min * C.dt;
        if (!(v_min > 0) || !std::isfinite(v_min))
            return fail(DCP_ERR_INVALID, "channel velocity chv(1)*dt must be positive and finite for every member (min %g)", v_min);
        double max_reach = 0, max_full = 0;
        for (int n = 0; n < C.nnode; ++n) {
            max_reach = std::max(max_reach, c->node_reach_dist[n] / v_min);
            max_full = std::max(max_full, (c->node_reach_dist[n] + c->node_down_dist[n]) / v_min);
        }
        if (max_full > 4.0e6) return fail(DCP_ERR_INVALID, "routing histogram would need %g bins (chv too small)", max_full);
        DevBatch W{};
        W.B = (int)Bn;
        W.Lcap = (int)std::ceil(max_reach) + 3;
        const int Tcap = (int)std::ceil(max_full) + 3;
        hist_len_max = std::max(hist_len_max, Tcap - 2);
        if (fused) {                             // rings: histogram span + the tile being filled + the tile being routed
            int R = 1;
            while (R < Tcap + DCP_RT_SLACK_STEPS) R <<= 1;
            W.R = R; W.Rmask = R - 1;
        } else if (use_res || tile >= C.nstep) { // whole series resident in the flow buffers: no ring
            W.R = C.nstep; W.Rmask = 0x7fffffff;
        } else {
            int R = 1;
            while (R < tile + Tcap + 1) R <<= 1;
            W.R = R; W.Rmask = R - 1;
        }
        const size_t ws_bytes = carve_batch(nullptr, C, W, check, aet, sw_nq, ring_slots, want_sig);
        if ((double)(ws_bytes + io_bytes) > mem_budget) {
            if (Bn <= 1) return fail(DCP_ERR_NOMEM, "one member needs %zu MB, device has %zu MB free",
                                     (ws_bytes + io_bytes) >> 20, free_b >> 20);
            Bmax = shrink(Bn);                  // retry this batch smaller
            continue;
        }
        CUDA_TRY(c->ws.reserve(ws_bytes));
        carve_batch(c->ws.p, C, W, check, aet, sw_nq, ring_slots, want_sig);
        if (want_sig) CUDA_TRY(cudaMemsetAsync(W.fdc, 0, (size_t)DCP_FDC_BINS * C.nriv * Bn * 4, s));
        W.q_user = fused ? d_q : nullptr;
        if (use_sw) CUDA_TRY(cudaMemsetAsync(W.bad_src, 0, (size_t)Bn * 4, s));

        tm.mark(T_INIT);
        dcp_launch_init(C, W, d_mcpar, check, s); ++n_launch;
        if (use_res) {
            ResPlan P = c->plan;
            P.n_groups = use_v2w ? 0 : (int)((Bn + P.Mc - 1) / P.Mc);
            const int grid = std::max(1, std::min(c->n_sm, P.n_groups));
            DevBuf dbg;
            if (getenv("DCP_RES_DEBUG")) { CUDA_TRY(dbg.reserve((size_t)grid * 20 * 8)); CUDA_TRY(cudaMemsetAsync(dbg.p, 0, (size_t)grid * 20 * 8, s)); P.dbg = (long long *)dbg.p; }
            Res2Plan Q = use_v2w ? c->plan2w : c->plan2;
            int grid2 = 0;
            if (use_v2) {
                const int Mv = Q.G * 2, teams = use_v2w ? 1 : 2;   // members per team group; teams per CTA
                Q.n_groups = (int)((Bn + Mv - 1) / Mv);
                grid2 = std::min(c->n_sm, (Q.n_groups + teams - 1) / teams);
                Q.dbg = P.dbg;
            }
            tm.mark(T_PHYS);
            if (use_v2w) CUDA_TRY(dcp_launch_physics_resident_v2w(C, W, Q, aet, grid2, s));
            else if (use_v2) CUDA_TRY(dcp_launch_physics_resident_v2(C, W, Q, aet, grid2, s));
            else CUDA_TRY(dcp_launch_physics_resident(C, W, P, fast, aet, grid, s));
            ++n_launch; ++n_phys;
            if (P.dbg) {           // per-warp busy cycles of the persistent kernel (tuning aid)
                std::vector<long long> hb((size_t)grid * 20);
                CUDA_TRY(cudaMemcpyAsync(hb.data(), dbg.p, hb.size() * 8, cudaMemcpyDeviceToHost, s));
                CUDA_TRY(cudaStreamSynchronize(s));
                for (int b = 0; b < std::min(grid, 3); ++b) {
                    fprintf(stderr, "[dcp resident dbg] cta %d busy Mcycles per warp:", b);
                    for (int w2 = 0; w2 < 20; ++w2) fprintf(stderr, " %.2f", hb[(size_t)b * 20 + w2] * 1e-6);
                    fprintf(stderr, "\n");
                }
                dbg.release();
            }
            tm.mark(T_ROUTE);
            if (!fused) { CUDA_TRY(dcp_launch_route(C, W, 0, C.nstep, d_q, s)); n_launch += C.qobs ? 3 : 2; }
        } else {
            for (int t0 = 0; t0 < C.nstep; t0 += tile) {
                const int t1 = std::min(C.nstep, t0 + tile);
                tm.mark(T_PHYS);
                if (use_sw) {
                    CUDA_TRY(dcp_launch_physics_stepwise(C, W, c->dplan, t0, t1, fast, aet, sw_dense, c->n_sm, s));
                    n_launch += (int64_t)(t1 - t0) * (sw_dense ? 4 : 3); n_phys += t1 - t0;
                } else { CUDA_TRY(dcp_launch_physics_streamed(C, W, t0, t1, fast, check, aet, s)); ++n_launch; ++n_phys; }
                tm.mark(T_ROUTE);
                CUDA_TRY(dcp_launch_route(C, W, t0, t1, d_q, s)); n_launch += C.qobs ? 3 : 2;
            }
        }
        tm.mark(T_FIN);
        dcp_launch_finalize(C, W, d_perf, d_tot, d_diag, d_status, d_sig, s); ++n_launch;
        tm.mark(T_D2H);
        if (!dev_io) {
            CUDA_TRY(cudaMemcpyAsync(mcpar + m0 * par_per_member, d_mcpar, par_per_member * Bn * 8, cudaMemcpyDeviceToHost, s));
            if (q_out) CUDA_TRY(cudaMemcpyAsync(q_out + (size_t)m0 * C.nriv * C.nstep, d_q, (size_t)C.nriv * C.nstep * Bn * 8, cudaMemcpyDeviceToHost, s));
            if (want_perf) CUDA_TRY(cudaMemcpyAsync(perf_out + (size_t)m0 * DCP_NMEASURE * C.nriv, d_perf, (size_t)DCP_NMEASURE * C.nriv * Bn * 8, cudaMemcpyDeviceToHost, s));
            if (want_sig) CUDA_TRY(cudaMemcpyAsync(o.sig_out + (size_t)m0 * DCP_NSIG * C.nriv, d_sig, (size_t)DCP_NSIG * C.nriv * Bn * 8, cudaMemcpyDeviceToHost, s));
            if (reach_totals) CUDA_TRY(cudaMemcpyAsync(reach_totals + (size_t)m0 * 3 * C.nriv, d_tot, (size_t)3 * C.nriv * Bn * 8, cudaMemcpyDeviceToHost, s));
            if (init_diag) CUDA_TRY(cudaMemcpyAsync(init_diag + (size_t)m0 * 3, d_diag, (size_t)3 * Bn * 8, cudaMemcpyDeviceToHost, s));
            if (status) CUDA_TRY(cudaMemcpyAsync(status + m0, d_status, (size_t)Bn * 4, cudaMemcpyDeviceToHost, s));
        }
        tm.mark(-1);
        CUDA_TRY(cudaStreamSynchronize(s));
        CUDA_TRY(cudaGetLastError());
        m0 += Bn;
    }
    double sums[T_NTAGS];
    tm.collect(sums, T_NTAGS);
    c->stats.ms_total = tm.total();
    c->stats.ms_h2d = sums[T_H2D];
    c->stats.ms_init = sums[T_INIT];
    c->stats.ms_physics = sums[T_PHYS];
    c->stats.ms_route = sums[T_ROUTE] + sums[T_FIN];
    c->stats.ms_d2h = sums[T_D2H];
    c->stats.n_physics_launches = n_phys;
    c->stats.n_launches = n_launch;
    c->stats.hist_len = hist_len_max;
    return DCP_OK;
}

int dcp_route_timeseries(dcp_catchment *c, int64_t n_series, const double *velocity, const double *inflow,
                         double *routed, int32_t *status, int32_t flags) {
    if (!c || !velocity || !inflow || !routed) return fail(DCP_ERR_INVALID, "null handle, velocity, inflow or routed");
    if (n_series < 1) return fail(DCP_ERR_INVALID, "n_series < 1");
    if (flags & ~DCP_OPT_DEVICE_BUFFERS) return fail(DCP_ERR_INVALID, "dcp_route_timeseries takes DCP_OPT_DEVICE_BUFFERS only");
    CUDA_TRY(cudaSetDevice(c->device));
    const bool dev_io = (flags & DCP_OPT_DEVICE_BUFFERS) != 0;
    DevCatchment C = c->C;
    C.qobs = nullptr;                           // no objective terms
    C.frac_sub_area = c->ones_riv;              // route_process_run_timeseries returns the plain sums (:685-714)
    cudaStream_t s = c->stream;
    c->stats = dcp_run_stats{};
    EventTimer tm(s);
    int64_t n_launch = 0;
    int hist_len_max = 0;
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    free_b += c->ws.cap + c->io.cap;
    const double mem_budget = 0.85 * (double)free_b;
    const size_t series_elems = (size_t)C.nriv * C.nstep;
    std::vector<double> v_host;
    const double *v_all = velocity;
    if (dev_io) {                               // the slowest velocity of a batch sizes its histograms
        v_host.resize((size_t)n_series);
        CUDA_TRY(cudaMemcpy(v_host.data(), velocity, (size_t)n_series * 8, cudaMemcpyDeviceToHost));
        v_all = v_host.data();
    }
    int64_t Bmax = std::min<int64_t>(n_series, 65536);
    for (int64_t m0 = 0; m0 < n_series;) {
        int64_t Bn = std::min<int64_t>(Bmax, n_series - m0);
        double v_min = v_all[m0];
        for (int64_t i = 0; i < Bn; ++i) {
            const double v = v_all[m0 + i];
            if (!(v > 0) || !std::isfinite(v)) return fail(DCP_ERR_INVALID, "velocity of series %lld must be positive and finite (%g)", (long long)(m0 + i), v);
            v_min = std::min(v_min, v);
        }
        double max_reach = 0, max_full = 0;
        for (int n = 0; n < C.nnode; ++n) {
            max_reach = std::max(max_reach, c->node_reach_dist[n] / v_min);
            max_full = std::max(max_full, (c->node_reach_dist[n] + c->node_down_dist[n]) / v_min);
        }
        if (max_full > 4.0e6) return fail(DCP_ERR_INVALID, "routing histogram would need %g bins (velocity too small)", max_full);
        DevBatch W{};
        W.B = (int)Bn;
        W.Lcap = (int)std::ceil(max_reach) + 3;
        hist_len_max = std::max(hist_len_max, (int)std::ceil(max_full) + 1);
        W.R = C.nstep; W.Rmask = 0x7fffffff;
        auto carve = [&](void *base) {
            Carver cv(base);
            W.dh = cv.take<double>((size_t)C.nnode * W.Lcap * Bn);
            W.hstart = cv.take<int32_t>((size_t)C.nnode * Bn);
            W.hbins = cv.take<int32_t>((size_t)C.nnode * Bn);
            W.qout = cv.take<double>(series_elems * Bn);
            W.y = cv.take<double>((size_t)C.nnode * C.nstep * Bn);
            W.bad = cv.take<int32_t>((size_t)C.nriv * Bn);
            W.status = cv.take<int32_t>(Bn);
            return cv.off + 256;
        };
        const size_t ws_bytes = carve(nullptr);
        const size_t io_bytes = dev_io ? 256 : (2 * series_elems * Bn + Bn) * 8 + Bn * 4 + 1024;
        if ((double)(ws_bytes + io_bytes) > mem_budget) {
            if (Bn <= 1) return fail(DCP_ERR_NOMEM, "one series needs %zu MB, device has %zu MB free", (ws_bytes + io_bytes) >> 20, free_b >> 20);
            Bmax = std::max<int64_t>(1, Bn / 2);
            continue;
        }
        CUDA_TRY(c->ws.reserve(ws_bytes));
        CUDA_TRY(c->io.reserve(io_bytes));
        carve(c->ws.p);
        const double *d_v = velocity + m0, *d_in = inflow + (size_t)m0 * series_elems;
        double *d_out = routed + (size_t)m0 * series_elems;
        int32_t *d_status = status ? status + m0 : nullptr;
        tm.mark(T_H2D);
        if (!dev_io) {
            Carver iocv(c->io.p);
            double *dv = iocv.take<double>(Bn), *din = iocv.take<double>(series_elems * Bn);
            d_out = iocv.take<double>(series_elems * Bn);
            d_status = iocv.take<int32_t>(Bn);
            CUDA_TRY(cudaMemcpyAsync(dv, velocity + m0, (size_t)Bn * 8, cudaMemcpyHostToDevice, s));
            CUDA_TRY(cudaMemcpyAsync(din, inflow + (size_t)m0 * series_elems, series_elems * Bn * 8, cudaMemcpyHostToDevice, s));
            d_v = dv; d_in = din;
        }
        tm.mark(T_INIT);
        dcp_launch_route_prepare(C, W, d_v, d_in, s); n_launch += 2;
        tm.mark(T_ROUTE);
        CUDA_TRY(dcp_launch_route(C, W, 0, C.nstep, d_out, s)); n_launch += 2;
        if (d_status) { dcp_launch_route_status(C, W, d_status, s); ++n_launch; }
        tm.mark(T_D2H);
        if (!dev_io) {
            CUDA_TRY(cudaMemcpyAsync(routed + (size_t)m0 * series_elems, d_out, series_elems * Bn * 8, cudaMemcpyDeviceToHost, s));
            if (status) CUDA_TRY(cudaMemcpyAsync(status + m0, d_status, (size_t)Bn * 4, cudaMemcpyDeviceToHost, s));
        }
        tm.mark(-1);
        CUDA_TRY(cudaStreamSynchronize(s));
        CUDA_TRY(cudaGetLastError());
        m0 += Bn;
    }
    double sums[T_NTAGS];
    tm.collect(sums, T_NTAGS);
    c->stats.ms_total = tm.total();
    c->stats.ms_h2d = sums[T_H2D];
    c->stats.ms_init = sums[T_INIT];
    c->stats.ms_route = sums[T_ROUTE];
    c->stats.ms_d2h = sums[T_D2H];
    c->stats.n_launches = n_launch;
    c->stats.hist_len = hist_len_max;
    return DCP_OK;
}

int dcp_select_behavioural(const double *perf, int64_t n_member, int32_t nriv, int32_t measure, int32_t gauge, double threshold,
                           int32_t higher_is_better, int64_t *index_out, double *weight_out, int64_t *n_selected, int32_t flags) {
    if (!perf || !index_out || !n_selected) return fail(DCP_ERR_INVALID, "dcp_select_behavioural: null argument");
    if (n_member < 1 || nriv < 1 || measure < 0 || measure >= DCP_NMEASURE || gauge < -1 || gauge >= nriv)
        return fail(DCP_ERR_INVALID, "dcp_select_behavioural: bad n_member / nriv / measure / gauge");
    if (flags & ~DCP_OPT_DEVICE_BUFFERS) return fail(DCP_ERR_INVALID, "dcp_select_behavioural takes DCP_OPT_DEVICE_BUFFERS only");
    if (dcp_device_count() < 1) return fail(DCP_ERR_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
    const size_t pb = (size_t)DCP_NMEASURE * nriv * (size_t)n_member * 8;
    const int nblk = dcp_select_blocks(n_member);
    DevBuf d_perf, d_cnt, d_off, d_idx, d_dist;
    const double *dp = perf;
    if (!(flags & DCP_OPT_DEVICE_BUFFERS)) {
        CUDA_TRY(d_perf.reserve(pb));
        CUDA_TRY(cudaMemcpy(d_perf.p, perf, pb, cudaMemcpyHostToDevice));
        dp = (const double *)d_perf.p;
    }
    CUDA_TRY(d_cnt.reserve((size_t)nblk * 4));
    CUDA_TRY(d_off.reserve((size_t)nblk * 8));
    CUDA_TRY(d_idx.reserve((size_t)n_member * 8));
    CUDA_TRY(d_dist.reserve((size_t)n_member * 8));
    dcp_launch_select_count(dp, n_member, nriv, measure, gauge, threshold, higher_is_better, (int32_t *)d_cnt.p, 0);
    std::vector<int32_t> cnt(nblk);
    CUDA_TRY(cudaMemcpy(cnt.data(), d_cnt.p, (size_t)nblk * 4, cudaMemcpyDeviceToHost));
    std::vector<int64_t> off(nblk);
    int64_t total = 0;
    for (int b = 0; b < nblk; ++b) { off[b] = total; total += cnt[b]; }
    CUDA_TRY(cudaMemcpy(d_off.p, off.data(), (size_t)nblk * 8, cudaMemcpyHostToDevice));
    dcp_launch_select_scatter(dp, n_member, nriv, measure, gauge, threshold, higher_is_better, (const int64_t *)d_off.p,
                              (int64_t *)d_idx.p, (double *)d_dist.p, 0);
    CUDA_TRY(cudaGetLastError());
    if (total > 0) CUDA_TRY(cudaMemcpy(index_out, d_idx.p, (size_t)total * 8, cudaMemcpyDeviceToHost));
    if (weight_out && total > 0) {
        CUDA_TRY(cudaMemcpy(weight_out, d_dist.p, (size_t)total * 8, cudaMemcpyDeviceToHost));
        double sum = 0.0;
        for (int64_t i = 0; i < total; ++i) sum = sum + weight_out[i];        // ascending member order
        for (int64_t i = 0; i < total; ++i) weight_out[i] = sum > 0.0 ? weight_out[i] / sum : 1.0 / (double)total;
    }
    *n_selected = total;
    return DCP_OK;
}

int dcp_debug_fastmath(int op, int64_t n, const double *x, double *y) {
    if (dcp_device_count() < 1) return fail(DCP_ERR_NO_DEVICE, "no CUDA device");
    if (op < 0 || op > 4 || n < 1 || !x || !y) return fail(DCP_ERR_INVALID, "bad argument");
    DevBuf bx, by, bt;
    CUDA_TRY(bx.reserve(n * 8));
    CUDA_TRY(by.reserve(n * 8));
    std::vector<double> tab(FM_EXP_TAB + 2 * FM_LOG_TAB);
    fm_build_tables(tab.data(), tab.data() + FM_EXP_TAB);
    CUDA_TRY(bt.reserve(tab.size() * 8));
    CUDA_TRY(cudaMemcpy(bt.p, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(bx.p, x, n * 8, cudaMemcpyHostToDevice));
    dcp_launch_debug_fastmath(op, n, (const double *)bx.p, (double *)by.p, (const double *)bt.p, 0);
    CUDA_TRY(cudaMemcpy(y, by.p, n * 8, cudaMemcpyDeviceToHost));
    return DCP_OK;
}

int dcp_measure_fp64_peak(double *tflops, double *ms_out) {
    if (dcp_device_count() < 1) return fail(DCP_ERR_NO_DEVICE, "no CUDA device");
    cudaDeviceProp prop;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    const int blocks = prop.multiProcessorCount * 8, iters = 1 << 16;
    double *buf = nullptr;
    CUDA_TRY(cudaMalloc(&buf, (size_t)blocks * 256 * 8));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a, 0);
        dcp_launch_fp64_peak(buf, iters, blocks, 0);
        cudaEventRecord(b, 0);
        CUDA_TRY(cudaEventSynchronize(b));
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        if (rep > 0) best = std::min(best, ms);
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(buf);
    const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * 256.0;
    if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return DCP_OK;
}

}  // extern "C"
