// fsv_capi.cu — host side of libfocalsv_cuda.so: the C ABI of include/focalsv_cuda.h.
//
// Replaces, for the alignment-DP hot path only, what the reference reaches through
//   minimap2 -> ksw_extd2_sse / ksw_extz2_sse   (software/hifiasm-0.16.1/ksw2.h:54-61)
// at DipPAV_variant_call.py:103-112, call_DUP_from_contigs.py:114-126 and
// align_ins2ref.py:64-71.  Ownership follows Correct.cpp:7673-7697 turned inside
// out: every buffer is caller-allocated, nothing owned is ever returned.
//
// There is no CPU path in this file: every result is produced by the CUDA kernels in
// fsv_fill_exact.cuh / fsv_fill_dpx.cuh / fsv_backtrack.cuh.
#include <algorithm>
#include <functional>
#include <cstdio>
#include <thread>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "fsv_backtrack.cuh"
#include "fsv_common.cuh"
#include "fsv_fill_dpx.cuh"
#include "fsv_fill_ew.cuh"
#include "fsv_fill_exact.cuh"
#include "fsv_peaks.cuh"
#include "fsv_editdist.cuh"
#include "fsv_signatures.cuh"

using namespace fsv;

// FSV_TRACE=1: host-side phase times of every batch call on stderr
#include <chrono>
struct HostTrace {
    bool on; std::chrono::steady_clock::time_point t;
    HostTrace() : on(getenv("FSV_TRACE") != nullptr), t(std::chrono::steady_clock::now()) {}
    void lap(const char* what) {
        if (!on) return;
        auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[fsv] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

struct fsv_ctx {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;            // copies, timing events
    cudaStream_t kstream[56] = {};            // one per concurrently running fill-kernel variant
    std::string last_error;
    fsv_stats stats{};
    // options
    int64_t tb_budget = 0;      // bytes of the traceback page pool (0 = auto: 70% of free memory, at most what the batch needs)
    int force_exact = 0;        // route every task to the general int8-exact kernel
    int force_excl = 0;         // experiment: every DPX task on the exclusive (one CTA per SM) launch
    int exact_smem_lanes = 4096;
    int64_t page_bytes = 32ll << 20;
    int64_t segment_min_diags = -1;  // tasks with at least this many antidiagonals are cut into segments (0 = off, -1 = auto:
                                     // those whose chain of antidiagonals would outlast 60 % of the batch's throughput time)
    int segment_warm_pct = 300;      // cold-start lead of a segment, in percent of the band width (+1024 antidiagonals); a boundary that does not verify costs one segment
    int segment_align_pages = 0;     // 0 = segments are multiples of 1024 antidiagonals (default), 1 = whole traceback pages
    int ew_kernel = 1;               // mainstream tasks run on the edge-warp kernel (fsv_fill_ew.cuh): 1 = those that need up to 4 main warps (bands up to
                                     // about 2 000; seven warps of 128 registers spill, so wider bands stay with fsv_fill_dpx_kernel), 2 = all of them, 0 = none
    int segment_auto_pct = 70;       // auto mode: tasks whose chain of antidiagonals outlasts this share of the batch's estimated time are segmented
    int segment_extz = 1;            // auto mode: 1 = extension (EXTZ_ONLY) tasks are segmented too, 0 = global tasks only
    int64_t segment_rows = 0;        // target antidiagonals per segment, rounded up to 1024 (0 = auto: 4 x or 2 x the warm-up)
    int lazy_min_pages = 16;    // DPX tasks with at least this many traceback pages take them as they advance (0 = all up front)
    int pool_stall_ms = 60000;  // lazy-pool watchdog
    int lazy_fill_pct = 65;     // admission of such a task waits while the projected peak of those running exceeds this share of the pool
    // scratch shared by the batches of this context (one batch runs at a time)
    uint8_t* d_pool = nullptr; size_t pool_cap = 0;
    uint8_t* d_ws = nullptr; size_t ws_cap = 0;
    int32_t* d_tables = nullptr; size_t tables_cap = 0;
    int32_t* d_free_stack = nullptr; size_t stack_cap = 0;
    int32_t* d_lazy = nullptr; size_t lazy_cap = 0;      // LazyState followed by the per-CTA slot index
    // device blocks of destroyed batches, kept for the next one (cudaMalloc / cudaFree of GB-sized blocks
    // cost tens of milliseconds per call and synchronise the device)
    std::vector<std::pair<void*, size_t>> dcache;
    std::vector<struct fsv_batch*> live;      // batches created on this context and not destroyed yet (fsv_destroy detaches them)
};

static void* dev_alloc(fsv_ctx* c, size_t bytes, size_t* got)
{
    int best = -1;
    for (size_t i = 0; i < c->dcache.size(); ++i) {
        const size_t sz = c->dcache[i].second;
        if (sz >= bytes && sz <= bytes + bytes / 2 + (1u << 20) && (best < 0 || sz < c->dcache[best].second)) best = (int)i;
    }
    if (best >= 0) {
        void* p = c->dcache[best].first; *got = c->dcache[best].second;
        c->dcache.erase(c->dcache.begin() + best);
        return p;
    }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {      // make room: drop what is cached, retry once
        cudaGetLastError();
        for (auto& e : c->dcache) cudaFree(e.first);
        c->dcache.clear();
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    }
    *got = bytes;
    return p;
}
static void dev_release(fsv_ctx* c, void* p, size_t bytes)
{
    if (!p) return;
    c->dcache.emplace_back(p, bytes);
    while (c->dcache.size() > 24) { cudaFree(c->dcache.front().first); c->dcache.erase(c->dcache.begin()); }
}

// one fill-kernel launch: a slice of the work list, largest task first
struct Launch { int kind /*0 general, 1 DPX*/, nw, with_tb, excl, begin, count, grid; int64_t table_off; };

// any base code outside A,C,G,T (0..3)?  8 bytes at a time.
static bool has_wildcard(const uint8_t* p, size_t n)
{
    size_t i = 0;
    uint64_t acc = 0;
    for (; i + 8 <= n; i += 8) { uint64_t x; memcpy(&x, p + i, 8); acc |= x; }
    uint8_t tail = 0;
    for (; i < n; ++i) tail |= p[i];
    return ((acc & 0xfcfcfcfcfcfcfcfcull) | (uint64_t)(tail & 0xfc)) != 0;
}

struct fsv_batch {
    fsv_ctx* ctx = nullptr;
    DevScoring sc{};
    bool dual = false;
    size_t n = 0;
    std::vector<DevTask> tasks;       // caller order
    std::vector<int32_t> work;        // task indices grouped per launch, largest first inside a group
    std::vector<Launch> launches;
    std::vector<uint8_t> is_dpx;      // per task
    int64_t cigar_cap_words = 0;
    int64_t ws_lanes = 0;
    int64_t pool_pages = 0;           // pages of the traceback pool this batch wants
    int64_t pages_total = 0;          // pages all its tasks need together
    int64_t cap_pages = 0;            // pages the pool may have at most (memory budget)
    int32_t max_pages_per_task = 1;
    int64_t tb_bytes_total = 0;       // traceback bytes a full run writes
    int64_t max_rows = 1;             // antidiagonals of the longest task that stores traceback
    // device
    uint8_t *d_q = nullptr, *d_t = nullptr;
    DevTask* d_tasks = nullptr;
    int32_t* d_work = nullptr;
    fsv_result* d_results = nullptr;
    int32_t* d_ctrl = nullptr;        // [0] overflow flag, [1] pool lock, [2] pool n_free; [8..] queue states (8 B each)
    unsigned long long* d_cursor = nullptr;
    uint32_t* d_cigar = nullptr;
    long long* d_timeline = nullptr;
    size_t sz_q = 0, sz_t = 0, sz_tasks = 0, sz_work = 0, sz_results = 0, sz_ctrl = 0, sz_cursor = 0, sz_cigar = 0, sz_timeline = 0;
    int state = 0;                    // 0 created, 1 run
    std::vector<uint8_t> is_seg, is_excl;   // per task: cut into segments / runs on the exclusive launch (fsv_batch_plan)
    // segmented tasks
    std::vector<DevSeg> segs;
    std::vector<SegTask> seg_tasks;
    int64_t seg_table_words = 0;          // page-table entries of the segmented tasks (filled on the device when a task is admitted)
    int64_t seg_rec_total = 0, seg_snap_words = 0;
    struct SegLaunch { int nw, begin, count; };
    std::vector<SegLaunch> seg_launches;  // slices of seg_work
    std::vector<int32_t> seg_work;        // segment indices per launch
    DevSeg* d_segs = nullptr; SegTask* d_seg_tasks = nullptr; int32_t* d_seg_tables = nullptr; int32_t* d_seg_work = nullptr;
    int32_t* d_seg_done = nullptr; int32_t* d_seg_foot = nullptr; int4* d_seg_rec = nullptr; uint32_t* d_seg_snap = nullptr;
    size_t sz_seg[8] = {0};
};

static const char* kErr[] = {"ok", "no CUDA device (libfocalsv_cuda has no CPU path)", "CUDA runtime error",
                             "invalid argument", "out of memory", "CIGAR arena too small",
                             "scoring outside the int8 range of ksw2", "batch used out of order"};

extern "C" const char* fsv_strerror(int code)
{
    int k = -code;
    if (k < 0 || k > 7) return "unknown error";
    return kErr[k];
}
extern "C" int fsv_abi_version(void) { return FSV_ABI_VERSION; }
extern "C" const char* fsv_last_error(const fsv_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

extern "C" int fsv_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

#define CK(ctx, call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char b_[512];                                                                          \
            snprintf(b_, sizeof b_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            (ctx)->last_error = b_;                                                                \
            cudaGetLastError();                                                                    \
            return e_ == cudaErrorMemoryAllocation ? FSV_ERR_NOMEM : FSV_ERR_CUDA;                 \
        }                                                                                          \
    } while (0)

extern "C" int fsv_init(int device, fsv_ctx** out)
{
    if (!out) return FSV_ERR_INVALID;
    *out = nullptr;
    int n = fsv_device_count();
    if (n <= 0) return FSV_ERR_NO_DEVICE;
    fsv_ctx* c = new (std::nothrow) fsv_ctx();
    if (!c) return FSV_ERR_NOMEM;
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= n) { delete c; return FSV_ERR_INVALID; }
    c->device = device;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete c; cudaGetLastError(); return FSV_ERR_CUDA;
    }
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; cudaGetLastError(); return FSV_ERR_CUDA; }
    for (auto& ks : c->kstream)
        if (cudaStreamCreateWithFlags(&ks, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); ks = nullptr; }
    *out = c;
    return FSV_OK;
}

static void free_batch_device(fsv_batch* b);
extern "C" void fsv_destroy(fsv_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    // batches that outlive their context are detached: their device blocks go back now, fsv_batch_destroy only
    // deletes the host object, every other call on them returns FSV_ERR_STATE
    for (fsv_batch* b : c->live) { free_batch_device(b); b->ctx = nullptr; b->state = -1; }
    c->live.clear();
    if (c->d_pool) cudaFree(c->d_pool);
    if (c->d_ws) cudaFree(c->d_ws);
    if (c->d_tables) cudaFree(c->d_tables);
    if (c->d_free_stack) cudaFree(c->d_free_stack);
    if (c->d_lazy) cudaFree(c->d_lazy);
    for (auto& e : c->dcache) cudaFree(e.first);
    for (auto ks : c->kstream) if (ks) cudaStreamDestroy(ks);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int fsv_get_stats(const fsv_ctx* c, fsv_stats* out)
{
    if (!c || !out) return FSV_ERR_INVALID;
    *out = c->stats;
    return FSV_OK;
}

extern "C" int fsv_set_option(fsv_ctx* c, const char* key, int64_t value)
{
    if (!c || !key) return FSV_ERR_INVALID;
    if (!strcmp(key, "traceback_budget_bytes")) { c->tb_budget = value; return FSV_OK; }
    if (!strcmp(key, "force_exact")) { c->force_exact = (int)value; return FSV_OK; }
    if (!strcmp(key, "force_excl")) { c->force_excl = (int)value; return FSV_OK; }
    if (!strcmp(key, "lazy_min_pages")) {      // 0 = off; a lazy task starts with two pages, so at least 3
        if (value < 0 || value > (1 << 20)) return FSV_ERR_INVALID;
        c->lazy_min_pages = value == 0 ? 0 : (int)std::max<int64_t>(value, 3); return FSV_OK;
    }
    if (!strcmp(key, "segment_min_diags")) { if (value < -1) return FSV_ERR_INVALID; c->segment_min_diags = value; return FSV_OK; }
    if (!strcmp(key, "segment_warm_pct")) { if (value < 50 || value > 2000) return FSV_ERR_INVALID; c->segment_warm_pct = (int)value; return FSV_OK; }
    if (!strcmp(key, "segment_slots") || !strcmp(key, "segment_pool_pct") || !strcmp(key, "segment_pool_pct_bound")) return FSV_OK;   // ABI 3 tunables of the static page slots: accepted, no effect
    if (!strcmp(key, "segment_auto_pct")) { if (value < 1 || value > 1000) return FSV_ERR_INVALID; c->segment_auto_pct = (int)value; return FSV_OK; }
    if (!strcmp(key, "segment_align_pages")) { c->segment_align_pages = value != 0; return FSV_OK; }
    if (!strcmp(key, "ew_kernel")) { if (value < 0 || value > 2) return FSV_ERR_INVALID; c->ew_kernel = (int)value; return FSV_OK; }
    if (!strcmp(key, "segment_extz")) { c->segment_extz = value != 0; return FSV_OK; }
    if (!strcmp(key, "segment_rows")) { if (value != 0 && value < 1024) return FSV_ERR_INVALID; c->segment_rows = value; return FSV_OK; }
    if (!strcmp(key, "pool_stall_ms")) { if (value < 100) return FSV_ERR_INVALID; c->pool_stall_ms = (int)value; return FSV_OK; }
    if (!strcmp(key, "lazy_fill_pct")) {
        if (value < 1 || value > 100) return FSV_ERR_INVALID;
        c->lazy_fill_pct = (int)value; return FSV_OK;
    }
    if (!strcmp(key, "traceback_page_bytes")) {
        if (value < (1 << 16) || (value & 255)) return FSV_ERR_INVALID;
        c->page_bytes = value; return FSV_OK;
    }
    if (!strcmp(key, "exact_smem_lanes")) {
        if (value < 256 || (value & (value - 1))) return FSV_ERR_INVALID;
        c->exact_smem_lanes = (int)value; return FSV_OK;
    }
    return FSV_ERR_INVALID;
}

// ---------------------------------------------------------------------------
extern "C" int64_t fsv_task_cells(int32_t qlen, int32_t tlen, int32_t w)
{
    if (qlen <= 0 || tlen <= 0) return 0;
    if (w < 0) w = tlen > qlen ? tlen : qlen;
    int64_t n = 0;
    for (int r = 0; r < qlen + tlen - 1; ++r) {
        int st0, en0;
        band_limits(r, qlen, tlen, w, st0, en0);
        if (st0 > en0) break;
        n += en0 - st0 + 1;
    }
    return n;
}

static inline int64_t cells_estimate(int qlen, int tlen, int w)
{
    if (qlen <= 0 || tlen <= 0) return 0;
    int64_t mn = std::min(qlen, tlen);
    if (w < 0) w = std::max(qlen, tlen);
    return (int64_t)(qlen + tlen - 1) * std::min<int64_t>(mn, (int64_t)w + 1);
}

extern "C" int fsv_lpt_bins(const fsv_task* tasks, size_t n, int n_bins, int32_t* bin_of)
{
    if ((!tasks && n) || !bin_of || n_bins <= 0) return FSV_ERR_INVALID;
    std::vector<int64_t> est(n);
    std::vector<size_t> idx(n);
    for (size_t i = 0; i < n; ++i) { est[i] = cells_estimate(tasks[i].qlen, tasks[i].tlen, tasks[i].w); idx[i] = i; }
    std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return est[a] > est[b]; });
    std::vector<int64_t> load((size_t)n_bins, 0);
    for (size_t k = 0; k < n; ++k) {
        int best = 0;
        for (int b = 1; b < n_bins; ++b) if (load[b] < load[best]) best = b;
        bin_of[idx[k]] = best;
        load[best] += est[idx[k]] + 1;
    }
    return FSV_OK;
}

// ---------------------------------------------------------------------------
static int build_scoring(const fsv_scoring* in, DevScoring* sc, bool* dual, int* all_reset_status, bool* all_reset)
{
    memset(sc, 0, sizeof(*sc));
    *all_reset = false; *all_reset_status = 0;
    int m = in->m;
    if (m > 5 || m * m > 32) return FSV_ERR_INVALID;
    *dual = in->q2 >= 0;
    int q = in->q, e = in->e, q2 = in->q2, e2 = in->e2;
    sc->m = m; sc->dual = *dual;
    if (m <= 0 || (*dual && m <= 1)) { *all_reset = true; return FSV_OK; }   // ksw2_extz2_sse.c:57
    if (*dual && q2 + e2 < q + e) { std::swap(q, q2); std::swap(e, e2); }
    sc->q = q; sc->e = e; sc->q2 = q2; sc->e2 = e2;
    memcpy(sc->mat, in->mat, (size_t)m * m);
    int max_sc = in->mat[0], min_sc = in->mat[1];
    for (int t = 1; t < m * m; ++t) { max_sc = std::max<int>(max_sc, in->mat[t]); min_sc = std::min<int>(min_sc, in->mat[t]); }
    if (-min_sc > 2 * (q + e)) { *all_reset = true; *all_reset_status = FSV_ERR_SCORING; return FSV_OK; }   // :82
    sc->sc_mch = (uint8_t)in->mat[0]; sc->sc_mis = (uint8_t)in->mat[1];
    int last = in->mat[m * m - 1];
    sc->sc_N = last == 0 ? (uint8_t)(-(*dual ? e2 : e)) : (uint8_t)last;      // :68
    if (*dual) {
        sc->max_sc_clamp = (uint8_t)in->mat[0];
        int lt = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
        if (q2 + e2 + lt * e2 > q + e + lt * e) ++lt;
        sc->long_thres = lt;
        sc->long_diff = lt * (e - e2) - (q2 - q) - e2;
        sc->e_drop = e2;
    } else {
        sc->max_sc_clamp = (uint8_t)(in->mat[0] + (q + e) * 2);               // :70
        sc->e_drop = e;
    }
    return FSV_OK;
}

static int64_t pow2_at_least(int64_t x) { int64_t p = 1; while (p < x) p <<= 1; return p; }

static void free_batch_device(fsv_batch* b)
{
    fsv_ctx* c = b->ctx;
    dev_release(c, b->d_q, b->sz_q); dev_release(c, b->d_t, b->sz_t); dev_release(c, b->d_tasks, b->sz_tasks);
    dev_release(c, b->d_work, b->sz_work); dev_release(c, b->d_results, b->sz_results); dev_release(c, b->d_ctrl, b->sz_ctrl);
    dev_release(c, b->d_cursor, b->sz_cursor); dev_release(c, b->d_cigar, b->sz_cigar); dev_release(c, b->d_timeline, b->sz_timeline);
    b->d_q = b->d_t = nullptr; b->d_tasks = nullptr; b->d_work = nullptr; b->d_results = nullptr; b->d_ctrl = nullptr;
    b->d_cursor = nullptr; b->d_cigar = nullptr; b->d_timeline = nullptr;
    dev_release(c, b->d_segs, b->sz_seg[0]); dev_release(c, b->d_seg_tasks, b->sz_seg[1]); dev_release(c, b->d_seg_tables, b->sz_seg[2]);
    dev_release(c, b->d_seg_work, b->sz_seg[3]); dev_release(c, b->d_seg_done, b->sz_seg[4]); dev_release(c, b->d_seg_foot, b->sz_seg[5]);
    dev_release(c, b->d_seg_rec, b->sz_seg[6]); dev_release(c, b->d_seg_snap, b->sz_seg[7]);
    b->d_segs = nullptr; b->d_seg_tasks = nullptr; b->d_seg_tables = nullptr; b->d_seg_work = nullptr; b->d_seg_done = nullptr;
    b->d_seg_foot = nullptr; b->d_seg_rec = nullptr; b->d_seg_snap = nullptr;
}

extern "C" void fsv_batch_destroy(fsv_batch* b)
{
    if (!b) return;
    if (b->ctx) {
        cudaSetDevice(b->ctx->device);
        free_batch_device(b);
        auto& lv = b->ctx->live;
        lv.erase(std::remove(lv.begin(), lv.end(), b), lv.end());
    }
    delete b;
}

static int exact_grid(fsv_ctx* c, bool dual, int n_tasks)
{
    const size_t smem = (size_t)c->exact_smem_lanes * (EXACT_NARR + 4);
    int per_sm = 0;
    cudaError_t e;
    if (dual) {
        cudaFuncSetAttribute(fsv_fill_exact_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fsv_fill_exact_kernel<true>, EXACT_THREADS, smem);
    } else {
        cudaFuncSetAttribute(fsv_fill_exact_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fsv_fill_exact_kernel<false>, EXACT_THREADS, smem);
    }
    if (e != cudaSuccess) { cudaGetLastError(); per_sm = 1; }
    if (per_sm < 1) per_sm = 1;
    return std::max(1, std::min(n_tasks, c->sm_count * per_sm));
}

extern "C" int fsv_batch_create(fsv_ctx* c, const fsv_scoring* scoring,
                                const uint8_t* qarena, size_t qbytes, const uint8_t* tarena, size_t tbytes,
                                const fsv_task* tasks, size_t n, fsv_batch** out)
{
    if (!c || !scoring || !out || (n && !tasks)) return FSV_ERR_INVALID;
    *out = nullptr;
    if (n > 0x7ffffff0u) return FSV_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    HostTrace tr;
    fsv_batch* b = new (std::nothrow) fsv_batch();
    if (!b) return FSV_ERR_NOMEM;
    b->ctx = c; b->n = n;
    bool all_reset; int reset_status;
    int rc = build_scoring(scoring, &b->sc, &b->dual, &reset_status, &all_reset);
    if (rc != FSV_OK) { delete b; return rc; }

    // ---- which tasks hold a base outside A,C,G,T (the DPX kernel packs bases in 2 bits): the one pass
    // over the caller's arenas, spread over host threads when they are large
    std::vector<uint8_t> wild(n, 0);
    if (!c->force_exact && n) {
        auto scan = [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i) {
                const fsv_task& t = tasks[i];
                if (t.qlen <= 0 || t.tlen <= 0 || t.q_off < 0 || t.t_off < 0 || (uint64_t)t.q_off + (uint64_t)t.qlen > qbytes ||
                    (uint64_t)t.t_off + (uint64_t)t.tlen > tbytes) continue;
                wild[i] = has_wildcard(qarena + t.q_off, (size_t)t.qlen) || has_wildcard(tarena + t.t_off, (size_t)t.tlen);
            }
        };
        unsigned nthr = (qbytes + tbytes) > (8u << 20) ? std::min(16u, std::max(1u, std::thread::hardware_concurrency())) : 1u;
        if (nthr <= 1) scan(0, n);
        else {
            std::vector<std::thread> th;
            for (unsigned k = 0; k < nthr; ++k) th.emplace_back(scan, n * k / nthr, n * (k + 1) / nthr);
            for (auto& t : th) t.join();
        }
    }
    tr.lap("create: wildcard scan");

    // ---- task table
    b->tasks.resize(n);
    b->is_dpx.assign(n, 0);
    int64_t cigar_words = 0, pages_total = 0;
    for (size_t i = 0; i < n; ++i) {
        const fsv_task& t = tasks[i];
        DevTask& d = b->tasks[i];
        memset(&d, 0, sizeof d);
        d.q_off = t.q_off; d.t_off = t.t_off; d.qlen = t.qlen; d.tlen = t.tlen;
        d.zdrop = t.zdrop; d.end_bonus = t.end_bonus; d.flag = t.flag; d.orig = (int32_t)i; d.tb_off = -1;
        d.kind = 1; d.rows_per_page = 1; d.seg_id = -1;
        if (all_reset || t.qlen <= 0 || t.tlen <= 0) { d.kind = 0; d.pad_ = all_reset ? reset_status : 0; continue; }
        if (t.q_off < 0 || t.t_off < 0 || (uint64_t)t.q_off + (uint64_t)t.qlen > qbytes ||
            (uint64_t)t.t_off + (uint64_t)t.tlen > tbytes || (int64_t)t.qlen + t.tlen > 0x7ffffff0) {
            c->last_error = "task " + std::to_string(i) + " points outside the sequence arenas";
            delete b; return FSV_ERR_INVALID;
        }
        int w = t.w < 0 ? std::max(t.qlen, t.tlen) : t.w;                         // :72
        d.w = w;
        int mn = std::min(t.qlen, t.tlen);
        int n_col = (std::min(mn, w + 1) + 15) / 16 + 1;                          // :75-76
        d.pitch = n_col * 16;
        d.cells_est = cells_estimate(t.qlen, t.tlen, w);
        if (!(t.flag & FSV_EZ_SCORE_ONLY)) {
            cigar_words += (int64_t)t.qlen + t.tlen + 2;
            if (d.pitch > c->page_bytes) {
                c->last_error = "task " + std::to_string(i) + ": one traceback row exceeds the page size";
                delete b; return FSV_ERR_NOMEM;
            }
            d.rows_per_page = (int32_t)(c->page_bytes / d.pitch);
            int64_t rows = (int64_t)t.qlen + t.tlen - 1;
            d.tb_pages = (int32_t)((rows + d.rows_per_page - 1) / d.rows_per_page);
            pages_total += d.tb_pages;
            b->max_pages_per_task = std::max(b->max_pages_per_task, d.tb_pages);
            b->tb_bytes_total += rows * d.pitch;
            b->max_rows = std::max(b->max_rows, rows);
        }
        if (!c->force_exact) {
            // the DPX kernel packs bases in 2 bits: a task with a wildcard base goes to the general kernel
            d.wild = wild[i];
            if (dpx_supports(b->sc, d, wild[i] != 0)) {
                b->is_dpx[i] = 1;
                d.nw = dpx_class_of(dpx_warps_needed(d));
                d.tb_mode = (t.flag & FSV_EZ_RIGHT) ? 0 : b->dual ? 4 : 2;      // right alignment stores ksw2's d itself
                // (kind 1: pad_ = main warps on the edge-warp kernel, 0 = that kernel does not take the task)
                if (c->ew_kernel && ew_supports(b->sc, d, wild[i] != 0)) {
                    const int nwm = dpx_class_of(ew_warps_needed(d));
                    if (nwm <= 4 || c->ew_kernel >= 2) d.pad_ = nwm;
                }
            }
        }
    }
    b->cigar_cap_words = cigar_words;
    tr.lap("create: task table");

    // ---- traceback page pool: what the batch needs, capped by the budget
    {
        int64_t budget = c->tb_budget;
        if (budget <= 0) {
            size_t fr = 0, tot = 0;
            CK(c, cudaMemGetInfo(&fr, &tot));
            fr += c->pool_cap;                       // the pool of a previous batch is reused
            budget = (int64_t)(fr * 0.88) - (int64_t)(cigar_words * 4);   // leave room for the CIGAR arena and the sequences
        }
        int64_t cap_pages = std::max<int64_t>(budget / c->page_bytes, 0);
        b->pool_pages = std::min(pages_total, cap_pages);
        b->pages_total = pages_total;
        b->cap_pages = cap_pages;
        if (b->max_pages_per_task > 1 || pages_total > 0)
            if (b->pool_pages < b->max_pages_per_task) {
                c->last_error = "the traceback of the largest task (" + std::to_string(b->max_pages_per_task) +
                                " pages) does not fit the pool (" + std::to_string(cap_pages) + " pages)";
                delete b; return FSV_ERR_NOMEM;
            }
    }

    // ---- work lists: one per kernel variant (general; DPX by warps-per-task class and with/without
    // traceback), each sorted largest-first (LPT within the device)
    std::vector<int32_t> ord(n);
    for (size_t i = 0; i < n; ++i) ord[i] = (int32_t)i;
    std::stable_sort(ord.begin(), ord.end(), [&](int32_t a, int32_t x) { return b->tasks[a].cells_est > b->tasks[x].cells_est; });
    // ---- segmented tasks (fsv_common.cuh, DevSeg): left-aligned CIGAR tasks whose chain of antidiagonals is long against the
    // batch are cut into segments that run on separate CTAs.  Their traceback comes from the dynamic pool like everybody
    // else's (all pages of a task at once, when its first segment starts), so how many are segmented is a question of work only:
    // a segment costs warm / seg_rows extra cells, a whole task that starts late keeps one CTA busy after the batch has ended.
    std::vector<uint8_t> is_seg(n, 0);
    if (c->segment_min_diags != 0) {
        // auto: a task is worth cutting up when its own chain of antidiagonals is long against the time the batch takes anyway.
        // That time is estimated per warp class with the SM-seconds model of the exclusive planning below: longest-first list
        // scheduling of the whole tasks on the class's CTA slots (so a uniform batch of a few waves is seen for what it is),
        // the work of the tasks already chosen spread over all slots.  Going down the eligible tasks by length, task i is
        // segmented while  chain_i > segment_auto_pct % of that estimate.
        int64_t min_diags = c->segment_min_diags;
        auto t_diag = [](int nw) { return nw <= 2 ? 2.0e-6 : nw == 4 ? 1.8e-6 : 1.55e-6; };
        auto occ_of = [](int nw) { return nw == 1 ? 12 : nw == 2 ? 6 : nw == 4 ? 3 : 2; };
        double W = 0;
        for (size_t i = 0; i < n; ++i) if (b->is_dpx[i]) W += (double)(b->tasks[i].qlen + b->tasks[i].tlen) * t_diag(b->tasks[i].nw) / occ_of(b->tasks[i].nw);
        const double t_thr = std::max(W / c->sm_count, 0.020);
        auto eligible = [&](size_t i) {
            const DevTask& d = b->tasks[i];
            return b->is_dpx[i] && d.tb_pages > 0 && !(d.flag & (FSV_EZ_RIGHT | FSV_EZ_SCORE_ONLY | FSV_EZ_APPROX_MAX)) && d.w >= 64 &&
                   d.tb_pages <= b->pool_pages && !(c->segment_min_diags < 0 && !c->segment_extz && (d.flag & FSV_EZ_EXTZ_ONLY));
        };
        std::vector<uint8_t> chosen(n, 0);
        if (min_diags < 0) {
            for (int nw : {8, 6, 4, 2, 1}) {
                std::vector<int32_t> cls;
                for (size_t i = 0; i < n; ++i) if (b->is_dpx[i] && b->tasks[i].nw == nw) cls.push_back((int32_t)i);
                if (cls.empty()) continue;
                std::stable_sort(cls.begin(), cls.end(), [&](int32_t a, int32_t x) { return b->tasks[a].qlen + b->tasks[a].tlen > b->tasks[x].qlen + b->tasks[x].tlen; });
                const int slots = c->sm_count * occ_of(nw);
                const double td = t_diag(nw);
                auto chain = [&](int32_t ti) { return (double)(b->tasks[(size_t)ti].qlen + b->tasks[(size_t)ti].tlen) * td; };
                double w_seg = 0;                    // slot-seconds of the tasks chosen so far, warm-up overhead included
                auto makespan = [&]() {
                    // list scheduling of the whole tasks (longest first) behind the divisible work of the chosen ones
                    std::vector<double> heap((size_t)slots, std::max(w_seg / slots, 0.0));
                    double m = heap[0];
                    const bool few_waves = cls.size() <= (size_t)8 * slots;
                    double w_whole = 0;
                    for (int32_t ti : cls) if (!chosen[(size_t)ti]) {
                        const double cch = chain(ti);
                        w_whole += cch;
                        if (!few_waves) { m = std::max(m, cch); continue; }
                        std::pop_heap(heap.begin(), heap.end(), std::greater<double>());
                        heap.back() += cch; m = std::max(m, heap.back());
                        std::push_heap(heap.begin(), heap.end(), std::greater<double>());
                    }
                    return std::max(std::max(m, (w_seg + w_whole) / slots), 0.020);
                };
                double M = makespan();
                int n_chosen = 0;
                bool fresh = true;
                for (int32_t ti : cls) {
                    if (!eligible((size_t)ti)) continue;
                    if (n_chosen >= 2048) break;
                    // the estimate falls as the longest chains turn into divisible work; it is recomputed (4 000 tasks through a heap) only when the
                    // last one says "stop": 20 ms of host time per call on cfg3 otherwise
                    if (chain(ti) <= c->segment_auto_pct * 0.01 * M) {
                        if (!fresh) { M = makespan(); fresh = true; }
                        if (chain(ti) <= c->segment_auto_pct * 0.01 * M) break;
                    }
                    chosen[(size_t)ti] = 1; ++n_chosen;
                    const DevTask& d = b->tasks[(size_t)ti];
                    const double warm = (double)c->segment_warm_pct * d.w / 100 + 1024, rows = std::max(4.0 * warm, 16384.0);
                    w_seg += chain(ti) * (1.0 + warm / rows);
                    fresh = false;
                }
            }
        }
        std::vector<int32_t> by_len;
        for (size_t i = 0; i < n; ++i)
            if (eligible(i) && (min_diags < 0 ? chosen[i] != 0 : (int64_t)b->tasks[i].qlen + b->tasks[i].tlen - 1 >= min_diags)) by_len.push_back((int32_t)i);
        std::stable_sort(by_len.begin(), by_len.end(), [&](int32_t a, int32_t x) { return b->tasks[a].qlen + b->tasks[a].tlen > b->tasks[x].qlen + b->tasks[x].tlen; });
        std::vector<std::vector<int32_t>> per_nw(9);
        struct SegPlan { int32_t ti; int64_t seg_rows, warm; int n_segs; };
        std::vector<SegPlan> plan;
        int total_segs = 0;
        // segments of 4 x warm rows cost 25 % more cells; if that leaves most SMs without a segment (a handful of long
        // tasks: cfg1's two contigs make 16), halve them: 2 x warm rows, 50 % more cells on SMs that would idle anyway
        for (int mult = 4; mult >= 2; mult -= 2) {
            plan.clear(); total_segs = 0;
            for (int32_t ti : by_len) {
                const DevTask& d = b->tasks[(size_t)ti];
                const int64_t n_diag = (int64_t)d.qlen + d.tlen - 1;
                const int64_t warm = (int64_t)c->segment_warm_pct * d.w / 100 + 1024;      // a flat cold start is bit-identical well within it; a boundary that is not is repaired
                const int64_t want = c->segment_rows > 0 ? c->segment_rows : std::max<int64_t>(mult * warm, 16384);
                const int64_t rpp = d.rows_per_page;
                const int64_t seg_rows = c->segment_align_pages ? std::max<int64_t>(1, (want + rpp - 1) / rpp) * rpp : (want + 1023) / 1024 * 1024;
                const int n_segs = (int)((n_diag + seg_rows - 1) / seg_rows);
                if (n_segs < 2 || seg_rows < 2 * warm) continue;
                plan.push_back({ti, seg_rows, warm, n_segs});
                total_segs += n_segs;
            }
            if (c->segment_rows > 0 || total_segs >= c->sm_count) break;
        }
        if (getenv("FSV_TRACE"))
            fprintf(stderr, "[fsv] segment plan: throughput time %.3f s, %s, chosen %zu -> %zu tasks in %d segments\n",
                    t_thr, min_diags < 0 ? "auto" : "fixed threshold", by_len.size(), plan.size(), total_segs);
        std::vector<int> n_in_class(9, 0);
        for (const SegPlan& sp : plan) {
            const int32_t ti = sp.ti;
            DevTask& d = b->tasks[(size_t)ti];
            const int64_t n_diag = (int64_t)d.qlen + d.tlen - 1;
            SegTask st{};
            st.rec_off = b->seg_rec_total; st.snap_off = b->seg_snap_words; st.table_off = (int32_t)b->seg_table_words;
            st.n_segs = sp.n_segs; st.first_seg = (int32_t)b->segs.size();
            st.ticket = n_in_class[(size_t)d.nw]++;                 // admission order inside its launch = queue order
            b->seg_rec_total += n_diag;
            b->seg_snap_words += (int64_t)2 * (sp.n_segs - 1) * SEG_SNAP_WORDS;
            b->seg_table_words += d.tb_pages;
            d.seg_id = (int32_t)b->seg_tasks.size();
            for (int k = 0; k < sp.n_segs; ++k) {
                DevSeg g{};
                g.task = ti; g.index = k; g.count = sp.n_segs;
                g.r_begin = (int32_t)(k * sp.seg_rows); g.r_end = (int32_t)std::min<int64_t>((k + 1) * sp.seg_rows, n_diag);
                g.r0 = (int32_t)std::max<int64_t>(0, g.r_begin - sp.warm);
                per_nw[(size_t)d.nw].push_back((int32_t)b->segs.size());
                b->segs.push_back(g);
            }
            b->seg_tasks.push_back(st);
            is_seg[(size_t)ti] = 1;
        }
        for (int nw = 8; nw >= 1; --nw) if (!per_nw[(size_t)nw].empty()) {
            b->seg_launches.push_back({nw, (int)b->seg_work.size(), (int)per_nw[(size_t)nw].size()});
            b->seg_work.insert(b->seg_work.end(), per_nw[(size_t)nw].begin(), per_nw[(size_t)nw].end());
        }
    }
    // Tasks long enough to decide the batch time by themselves get an SM each ("exclusive" launch): nothing can
    // shorten a task's chain of antidiagonals, but a CTA that has its SM to itself steps through it faster
    // (measured: 1.2-1.3 us per antidiagonal alone, 1.55-2.0 us when the SM is full).  Planning model, in
    // SM-seconds: a shared SM holds occ(nw) tasks at t_shared(nw) per antidiagonal; the k longest tasks are made
    // exclusive for the k that minimises  max(rest of the work on the other SMs, longest shared task, longest
    // exclusive task).
    std::vector<uint8_t> is_excl(n, 0);
    {
        auto t_shared = [](int nw) { return nw <= 2 ? 2.0e-6 : nw == 4 ? 1.8e-6 : 1.55e-6; };
        auto t_solo = [](int nw) { return nw >= 6 ? 1.3e-6 : 1.2e-6; };
        auto occ = [](int nw) { return nw == 1 ? 12.0 : nw == 2 ? 6.0 : nw == 4 ? 3.0 : 2.0; };
        std::vector<int> dpx;                       // DPX tasks, longest chain of antidiagonals first
        for (size_t k = 0; k < n; ++k) if (b->is_dpx[ord[k]] && !is_seg[(size_t)ord[k]]) dpx.push_back(ord[k]);
        std::stable_sort(dpx.begin(), dpx.end(), [&](int a, int x) { return b->tasks[a].qlen + b->tasks[a].tlen > b->tasks[x].qlen + b->tasks[x].tlen; });
        const int S = c->sm_count, m = (int)dpx.size();
        auto nd = [&](int i) { return (double)(b->tasks[dpx[(size_t)i]].qlen + b->tasks[dpx[(size_t)i]].tlen); };
        double W = 0;
        for (int i = 0; i < m; ++i) W += nd(i) * t_shared(b->tasks[dpx[(size_t)i]].nw) / occ(b->tasks[dpx[(size_t)i]].nw);
        int best_k = 0;
        double best_T = 0, w_excl = 0, solo_max = 0;
        for (int k = 0; k <= std::min(m, S); ++k) {
            if (k > 0) {
                const int nw = b->tasks[dpx[(size_t)k - 1]].nw;
                w_excl += nd(k - 1) * t_shared(nw) / occ(nw);
                solo_max = std::max(solo_max, nd(k - 1) * t_solo(nw));
            }
            if (k == S && k < m) break;             // the other tasks need an SM too
            const double rest = k < m ? std::max((W - w_excl) / (S - k), nd(k) * t_shared(b->tasks[dpx[(size_t)k]].nw)) : 0.0;
            const double T = std::max(rest, solo_max);
            if (k == 0 || T < best_T * 0.97) { best_T = T; best_k = k; }      // 3 % hysteresis: do not reserve SMs for nothing
        }
        if (c->force_excl) best_k = m;
        for (int i = 0; i < best_k; ++i) is_excl[(size_t)dpx[(size_t)i]] = 1;
    }
    int64_t ws_need = 0, table_off = 0;
    auto add_launch = [&](int kind, int nw, int with_tb, int excl) {
        Launch L{kind, nw, with_tb, excl, (int)b->work.size(), 0, 0, 0};
        for (size_t k = 0; k < n; ++k) {
            const int ti = ord[k];
            const DevTask& d = b->tasks[ti];
            if (is_seg[(size_t)ti]) continue;
            if (kind == 0) { if (b->is_dpx[ti]) continue; }
            else {
                // with_tb: 0 score only, 1 traceback (ties left), 2 traceback (ties right), 3 score only + approximate maximum, 4 = 1 on the edge-warp kernel
                const int mode = (d.flag & FSV_EZ_SCORE_ONLY) ? ((d.flag & FSV_EZ_APPROX_MAX) ? 3 : 0) : (d.flag & FSV_EZ_RIGHT) ? 2 : d.pad_ > 0 ? 4 : 1;
                if (!b->is_dpx[ti] || (mode == 4 ? d.pad_ : d.nw) != nw || mode != with_tb || (int)is_excl[ti] != excl) continue;
            }
            b->work.push_back(ti);
            if (kind == 0 && d.kind == 1 && d.pitch + 96 > c->exact_smem_lanes) ws_need = std::max<int64_t>(ws_need, d.pitch + 96);
        }
        L.count = (int)b->work.size() - L.begin;
        if (!L.count) return;
        L.grid = kind == 0 ? exact_grid(c, b->dual, L.count) : excl ? std::min(L.count, c->sm_count)
                 : with_tb == 4 ? (b->dual ? ew_grid_1(c->sm_count, nw, L.count) : ew_grid_0(c->sm_count, nw, L.count)) : dpx_grid(c->sm_count, b->dual, with_tb, nw, L.count);
        L.table_off = table_off;
        table_off += (int64_t)L.grid * b->max_pages_per_task;
        b->launches.push_back(L);
    };
    static const int kClasses[5] = {8, 6, 4, 2, 1};
    for (int cls : kClasses) for (int with_tb = 4; with_tb >= 0; --with_tb) add_launch(1, cls, with_tb, 1);
    for (int cls : kClasses) for (int with_tb = 4; with_tb >= 0; --with_tb) add_launch(1, cls, with_tb, 0);
    add_launch(0, 0, 0, 0);
    tr.lap("create: sort + work lists");
    b->ws_lanes = ws_need ? pow2_at_least(ws_need) : 0;
    if (b->launches.size() > 48) { delete b; return FSV_ERR_INVALID; }

    // ---- device buffers + H2D
    auto fail = [&](int code) { free_batch_device(b); delete b; return code; };
#define CKB(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            c->last_error = std::string(#call) + " -> " + cudaGetErrorString(e_);                  \
            cudaGetLastError();                                                                    \
            return fail(e_ == cudaErrorMemoryAllocation ? FSV_ERR_NOMEM : FSV_ERR_CUDA);           \
        }                                                                                          \
    } while (0)
#define DEV(ptr, szf, bytes)                                                                        \
    do {                                                                                           \
        void* p_ = dev_alloc(c, (bytes), &b->szf);                                                 \
        if (!p_) { c->last_error = "device allocation of " + std::to_string((size_t)(bytes)) + " bytes failed"; return fail(FSV_ERR_NOMEM); } \
        b->ptr = reinterpret_cast<decltype(b->ptr)>(p_);                                           \
    } while (0)
    DEV(d_q, sz_q, qbytes + 64);
    DEV(d_t, sz_t, tbytes + 64);
    DEV(d_tasks, sz_tasks, (n + 1) * sizeof(DevTask));
    DEV(d_work, sz_work, (n + 1) * 4);
    DEV(d_results, sz_results, (n + 1) * sizeof(fsv_result));
    DEV(d_ctrl, sz_ctrl, 512);
    DEV(d_cursor, sz_cursor, 64);
    DEV(d_cigar, sz_cigar, (size_t)(b->cigar_cap_words + 4) * 4);
    DEV(d_timeline, sz_timeline, (n + 1) * 16);
    if (!b->segs.empty()) {
        DEV(d_segs, sz_seg[0], b->segs.size() * sizeof(DevSeg));
        DEV(d_seg_tasks, sz_seg[1], b->seg_tasks.size() * sizeof(SegTask));
        DEV(d_seg_tables, sz_seg[2], (size_t)b->seg_table_words * 4 + 16);
        DEV(d_seg_work, sz_seg[3], b->seg_work.size() * 4 + 16);
        DEV(d_seg_done, sz_seg[4], b->seg_tasks.size() * 12 + 64);      // done counters, cancel flags, admitted flags, ticket counters (one per launch)
        DEV(d_seg_foot, sz_seg[5], b->segs.size() * 4 + 16);
        DEV(d_seg_rec, sz_seg[6], (size_t)b->seg_rec_total * sizeof(int4) + 16);
        DEV(d_seg_snap, sz_seg[7], (size_t)b->seg_snap_words * 4 + 16);
    }
#undef DEV
    tr.lap("create: cudaMalloc");
    CKB(cudaMemsetAsync(b->d_timeline, 0, (n + 1) * 16, c->stream));
    if (qbytes) CKB(cudaMemcpyAsync(b->d_q, qarena, qbytes, cudaMemcpyHostToDevice, c->stream));
    if (tbytes) CKB(cudaMemcpyAsync(b->d_t, tarena, tbytes, cudaMemcpyHostToDevice, c->stream));
    if (n) {
        CKB(cudaMemcpyAsync(b->d_tasks, b->tasks.data(), n * sizeof(DevTask), cudaMemcpyHostToDevice, c->stream));
        if (!b->work.empty()) CKB(cudaMemcpyAsync(b->d_work, b->work.data(), b->work.size() * 4, cudaMemcpyHostToDevice, c->stream));
        if (!b->segs.empty()) {
            CKB(cudaMemcpyAsync(b->d_segs, b->segs.data(), b->segs.size() * sizeof(DevSeg), cudaMemcpyHostToDevice, c->stream));
            CKB(cudaMemcpyAsync(b->d_seg_tasks, b->seg_tasks.data(), b->seg_tasks.size() * sizeof(SegTask), cudaMemcpyHostToDevice, c->stream));
            CKB(cudaMemcpyAsync(b->d_seg_work, b->seg_work.data(), b->seg_work.size() * 4, cudaMemcpyHostToDevice, c->stream));
        }
    }
    CKB(cudaStreamSynchronize(c->stream));
    tr.lap("create: H2D");
#undef CKB
    c->stats.h2d_bytes += (int64_t)(qbytes + tbytes + n * (sizeof(DevTask) + 4));
    b->is_seg.swap(is_seg); b->is_excl.swap(is_excl);
    c->live.push_back(b);
    *out = b;
    return FSV_OK;
}

template <typename T>
static int grow(fsv_ctx* c, T** p, size_t* cap, size_t bytes)
{
    if (bytes <= *cap) return FSV_OK;
    if (*p) { cudaFree(*p); *p = nullptr; *cap = 0; }
    CK(c, cudaMalloc(p, bytes + 256));
    *cap = bytes;
    return FSV_OK;
}

extern "C" int fsv_batch_run(fsv_batch* b)
{
    if (!b) return FSV_ERR_INVALID;
    fsv_ctx* c = b->ctx;
    if (!c) return FSV_ERR_STATE;             // its context was destroyed
    HostTrace tr;
    CK(c, cudaSetDevice(c->device));
    // ---- scratch: page pool, free stack, page tables, general-kernel windows
    int64_t tables = 0, ws_bytes = 0, n_slots = 0;
    for (auto& L : b->launches) {
        n_slots += L.grid;
        tables = std::max<int64_t>(tables, L.table_off + (int64_t)L.grid * b->max_pages_per_task);
        if (L.kind == 0 && b->ws_lanes) ws_bytes = (int64_t)L.grid * b->ws_lanes * (EXACT_NARR + 4);
    }
    int rc;
    if ((rc = grow(c, &c->d_pool, &c->pool_cap, (size_t)(b->pool_pages * c->page_bytes))) != FSV_OK) return rc;
    if ((rc = grow(c, &c->d_free_stack, &c->stack_cap, (size_t)(b->pool_pages + 1) * 4)) != FSV_OK) return rc;
    if ((rc = grow(c, &c->d_tables, &c->tables_cap, (size_t)(tables + 1) * 4)) != FSV_OK) return rc;
    if ((rc = grow(c, &c->d_ws, &c->ws_cap, (size_t)ws_bytes)) != FSV_OK) return rc;
    const size_t lazy_bytes = sizeof(LazyState) + (size_t)(n_slots + 1) * 4;
    if ((rc = grow(c, &c->d_lazy, &c->lazy_cap, lazy_bytes)) != FSV_OK) return rc;
    CK(c, cudaMemsetAsync(c->d_lazy, 0, sizeof(LazyState), c->stream));
    CK(c, cudaMemsetAsync(reinterpret_cast<uint8_t*>(c->d_lazy) + sizeof(LazyState), 0xff, (size_t)(n_slots + 1) * 4, c->stream));
    {   // control block: overflow flag, pool lock, free count, queue states; free stack = every page
        const int64_t dyn_pages = b->pool_pages;
        std::vector<int32_t> stack((size_t)b->pool_pages + 1);
        for (int64_t i = 0; i < dyn_pages; ++i) stack[(size_t)i] = (int32_t)i;
        CK(c, cudaMemcpyAsync(c->d_free_stack, stack.data(), (size_t)(b->pool_pages + 1) * 4, cudaMemcpyHostToDevice, c->stream));
        int32_t ctrl[128] = {0};
        ctrl[2] = (int32_t)dyn_pages;
        for (size_t i = 0; i < b->launches.size(); ++i) {
            unsigned long long st = (unsigned long long)(unsigned)b->launches[i].count;    // head 0, tail count
            memcpy(&ctrl[8 + 2 * i], &st, 8);
        }
        for (size_t i = 0; i < b->seg_launches.size(); ++i) {
            unsigned long long st = (unsigned long long)(unsigned)b->seg_launches[i].count;
            memcpy(&ctrl[8 + 2 * (b->launches.size() + i)], &st, 8);
        }
        if (!b->segs.empty()) {
            CK(c, cudaMemsetAsync(b->d_seg_done, 0, b->seg_tasks.size() * 12 + 64, c->stream));
            CK(c, cudaMemsetAsync(b->d_seg_tables, 0xff, (size_t)b->seg_table_words * 4, c->stream));
            CK(c, cudaMemsetAsync(b->d_seg_foot, 0xff, b->segs.size() * 4, c->stream));
            CK(c, cudaMemsetAsync(b->d_seg_snap, 0, (size_t)b->seg_snap_words * 4, c->stream));
        }
        CK(c, cudaMemcpyAsync(b->d_ctrl, ctrl, sizeof ctrl, cudaMemcpyHostToDevice, c->stream));
        CK(c, cudaMemsetAsync(b->d_cursor, 0, 64, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));      // `stack` and `ctrl` are stack/heap temporaries
    }
    tr.lap("run: scratch + control block");
    RunCtx R{};
    R.qarena = b->d_q; R.tarena = b->d_t; R.tasks = b->d_tasks; R.results = b->d_results;
    R.pool.base = c->d_pool; R.pool.page_bytes = c->page_bytes; R.pool.n_pages = (int32_t)b->pool_pages;
    R.pool.free_stack = c->d_free_stack; R.pool.n_free = b->d_ctrl + 2; R.pool.lock = b->d_ctrl + 1; R.pool.progress = b->d_ctrl + 3; R.pool.gate = b->d_ctrl + 4; R.pool.reserve = b->d_ctrl + 5;
    // lazy growth only pays (and only costs) when the batch's traceback does not fit the pool at once
    const bool lazy_on = c->lazy_min_pages > 0 && b->pool_pages < b->pages_total;      // (static pages of segmented tasks count on both sides)
    R.pool.lazy = lazy_on ? reinterpret_cast<LazyState*>(c->d_lazy) : nullptr;
    R.pool.slot_idx = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(c->d_lazy) + sizeof(LazyState));
    R.pool.lazy_min_pages = lazy_on ? c->lazy_min_pages : 0;
    R.pool.lazy_fill = (float)c->lazy_fill_pct * 0.01f;
    R.pool.stall_ms = c->pool_stall_ms;
    R.pool.bucket_rows = (int32_t)(b->max_rows / LAZY_BUCKETS + 1);
    R.max_pages_per_task = b->max_pages_per_task;
    R.cigar = b->d_cigar; R.cigar_cursor = b->d_cursor; R.cigar_cap = b->cigar_cap_words; R.overflow = b->d_ctrl + 0;
    R.timeline = b->d_timeline;
    R.sc = b->sc;
    R.segs = b->d_segs; R.seg_tasks = b->d_seg_tasks; R.seg_rec = b->d_seg_rec; R.seg_snap = b->d_seg_snap;
    R.seg_tables = b->d_seg_tables; R.seg_done = b->d_seg_done; R.seg_cancel = b->d_seg_done + b->seg_tasks.size(); R.seg_admitted = b->d_seg_done + 2 * b->seg_tasks.size(); R.seg_ticket = b->d_seg_done + 3 * b->seg_tasks.size(); R.seg_foot = b->d_seg_foot;

    // events of this run, destroyed on every way out
    struct Events {
        cudaEvent_t e0 = nullptr, e1 = nullptr; std::vector<cudaEvent_t> done;
        ~Events() { for (auto e : done) if (e) cudaEventDestroy(e); if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
    } ev;
    ev.done.assign(b->launches.size() + b->seg_launches.size(), nullptr);
    cudaEvent_t& e0 = ev.e0; cudaEvent_t& e1 = ev.e1; std::vector<cudaEvent_t>& done = ev.done;
    CK(c, cudaEventCreate(&e0)); CK(c, cudaEventCreate(&e1));
    CK(c, cudaEventRecord(e0, c->stream));
    // every kernel variant runs concurrently on its own stream; they share the page pool, so the long
    // tasks of one class overlap the short tasks of all the others
    int32_t slot_base = 0;
    // order: exclusive launches (their CTAs need EMPTY SMs: behind the segments they would wait for a whole SM to
    // drain), then the segments of the long tasks (what the batch waits for), then everything else
    auto launch_segments = [&]() -> int {
    for (size_t i = 0; i < b->seg_launches.size(); ++i) {
        const auto& L = b->seg_launches[i];
        const size_t qi = b->launches.size() + i;
        cudaStream_t ks = c->kstream[qi] ? c->kstream[qi] : c->stream;
        CK(c, cudaStreamWaitEvent(ks, e0, 0));
        TaskQueue Q{reinterpret_cast<unsigned long long*>(b->d_ctrl + 8 + 2 * qi), b->d_seg_work + L.begin};
        R.page_tables = c->d_tables; R.slot_base = 0;
        DpxParams D{R, Q, DpxK{}, (int32_t)i};
        rc = b->dual ? dpx_launch_seg_1(ks, c->sm_count, L.nw, L.count, D, &c->last_error) : dpx_launch_seg_0(ks, c->sm_count, L.nw, L.count, D, &c->last_error);
        if (rc != FSV_OK) return rc;
        c->stats.fill_launches++;
        CK(c, cudaEventCreateWithFlags(&done[qi], cudaEventDisableTiming));
        CK(c, cudaEventRecord(done[qi], ks));
        CK(c, cudaStreamWaitEvent(c->stream, done[qi], 0));
    }
    return FSV_OK;
    };
    auto launch_one = [&](size_t i) -> int {
        const Launch& L = b->launches[i];
        cudaStream_t ks = c->kstream[i] ? c->kstream[i] : c->stream;
        CK(c, cudaStreamWaitEvent(ks, e0, 0));
        TaskQueue Q{reinterpret_cast<unsigned long long*>(b->d_ctrl + 8 + 2 * i), b->d_work + L.begin};
        R.page_tables = c->d_tables + L.table_off;
        R.slot_base = slot_base; slot_base += L.grid;
        if (L.kind == 1) {
            DpxParams D{R, Q, DpxK{}, 0};
            rc = L.with_tb == 4 ? (b->dual ? ew_launch_1(ks, L.nw, L.grid, L.excl != 0, D, &c->last_error) : ew_launch_0(ks, L.nw, L.grid, L.excl != 0, D, &c->last_error))
                                : dpx_launch(ks, b->dual, L.with_tb, L.nw, L.grid, L.excl != 0, D, &c->last_error);
            if (rc != FSV_OK) return rc;
        } else {
            FillParams P{R, Q, c->d_ws, b->ws_lanes, c->exact_smem_lanes};
            const size_t smem = (size_t)c->exact_smem_lanes * (EXACT_NARR + 4);
            if (b->dual) fsv_fill_exact_kernel<true><<<L.grid, EXACT_THREADS, smem, ks>>>(P);
            else fsv_fill_exact_kernel<false><<<L.grid, EXACT_THREADS, smem, ks>>>(P);
            CK(c, cudaGetLastError());
        }
        c->stats.fill_launches++;
        CK(c, cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
        CK(c, cudaEventRecord(done[i], ks));
        CK(c, cudaStreamWaitEvent(c->stream, done[i], 0));
        return FSV_OK;
    };
    for (size_t i = 0; i < b->launches.size(); ++i)
        if (b->launches[i].kind == 1 && b->launches[i].excl && (rc = launch_one(i)) != FSV_OK) return rc;
    if ((rc = launch_segments()) != FSV_OK) return rc;
    for (size_t i = 0; i < b->launches.size(); ++i)
        if (!(b->launches[i].kind == 1 && b->launches[i].excl) && (rc = launch_one(i)) != FSV_OK) return rc;
    CK(c, cudaEventRecord(e1, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    tr.lap("run: kernels");
    float ms = 0;
    CK(c, cudaEventElapsedTime(&ms, e0, e1));
    c->stats.total_ms = ms; c->stats.fill_ms = ms; c->stats.backtrack_ms = 0;   // the CIGAR walk runs inside the fill kernels
    c->stats.traceback_bytes = b->tb_bytes_total;
    c->stats.segmented_tasks = (int64_t)b->seg_tasks.size();
    c->stats.segment_fallbacks = 0;
    if (!b->seg_tasks.empty()) {
        std::vector<int32_t> sd(b->seg_tasks.size());
        CK(c, cudaMemcpy(sd.data(), b->d_seg_done, sd.size() * 4, cudaMemcpyDeviceToHost));
        for (int32_t v : sd) c->stats.segment_fallbacks += v >> 20;      // repaired segments
    }
    c->stats.tasks += (int64_t)b->n;
    int64_t ne = 0;
    for (size_t i = 0; i < b->n; ++i) if (!b->is_dpx[i]) ++ne;
    c->stats.exact_path_tasks += ne;
    b->state = 1;
    return FSV_OK;
}

extern "C" int fsv_batch_fetch(fsv_batch* b, fsv_result* out, uint32_t* cigar, size_t cigar_cap, size_t* cigar_used)
{
    if (!b || (!out && b->n)) return FSV_ERR_INVALID;
    fsv_ctx* c = b->ctx;
    if (!c || b->state != 1) return FSV_ERR_STATE;
    CK(c, cudaSetDevice(c->device));
    int64_t used = 0;
    if (b->n) CK(c, cudaMemcpyAsync(out, b->d_results, b->n * sizeof(fsv_result), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpyAsync(&used, b->d_cursor, 8, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    if (cigar_used) *cigar_used = (size_t)used;
    int64_t cells = 0;
    for (size_t i = 0; i < b->n; ++i) cells += out[i].cells;
    c->stats.cells += cells;
    c->stats.d2h_bytes += (int64_t)(b->n * sizeof(fsv_result) + 8);
    if ((size_t)used > cigar_cap || (used && !cigar)) return FSV_ERR_CIGAR_CAP;
    if (used) {
        CK(c, cudaMemcpyAsync(cigar, b->d_cigar, (size_t)used * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
        c->stats.d2h_bytes += used * 4;
    }
    return FSV_OK;
}

extern "C" int fsv_edit_distance_batch(fsv_ctx* c, const uint8_t* a_arena, size_t a_bytes, const uint8_t* b_arena, size_t b_bytes,
                                       const fsv_pair* pairs, size_t n, int32_t* dist)
{
    if (!c || (n && (!pairs || !dist)) || n > 0x7ffffff0u) return FSV_ERR_INVALID;
    if (!n) return FSV_OK;
    CK(c, cudaSetDevice(c->device));
    // ---- bounds, the symbol remap (at most 8 distinct byte values) and the work order
    int remap[256]; bool seen[256] = {false};
    int64_t max_b = 1;
    for (size_t i = 0; i < n; ++i) {
        const fsv_pair& p = pairs[i];
        if (p.a_len < 0 || p.b_len < 0 || p.a_off < 0 || p.b_off < 0 || (uint64_t)p.a_off + (uint64_t)p.a_len > a_bytes ||
            (uint64_t)p.b_off + (uint64_t)p.b_len > b_bytes) { c->last_error = "pair " + std::to_string(i) + " points outside the arenas"; return FSV_ERR_INVALID; }
        if (p.a_len && p.b_len) {
            for (int k = 0; k < p.a_len; ++k) seen[a_arena[p.a_off + k]] = true;
            for (int k = 0; k < p.b_len; ++k) seen[b_arena[p.b_off + k]] = true;
            max_b = std::max<int64_t>(max_b, p.b_len);
        }
    }
    int n_sym = 0;
    for (int v = 0; v < 256; ++v) { remap[v] = 0; if (seen[v]) remap[v] = n_sym++; }
    if (n_sym > ED_MAX_SYMBOLS) { c->last_error = "more than 8 distinct byte values in one edit-distance call"; return FSV_ERR_INVALID; }
    std::vector<uint8_t> ra(a_bytes + 1), rb(b_bytes + 1);
    for (size_t i = 0; i < a_bytes; ++i) ra[i] = (uint8_t)remap[a_arena[i]];
    for (size_t i = 0; i < b_bytes; ++i) rb[i] = (uint8_t)remap[b_arena[i]];
    std::vector<int32_t> order(n);
    for (size_t i = 0; i < n; ++i) order[i] = (int32_t)i;
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
        return (int64_t)pairs[x].a_len * pairs[x].b_len > (int64_t)pairs[y].a_len * pairs[y].b_len; });
    // ---- device buffers (cached blocks of the context)
    const int warps_per_cta = 4;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((n + warps_per_cta - 1) / warps_per_cta, (size_t)c->sm_count * 4));
    const int64_t pitch = (max_b + 15) / 16 * 16;
    size_t sz[7] = {0};
    uint8_t* d_a = (uint8_t*)dev_alloc(c, a_bytes + 16, &sz[0]);
    uint8_t* d_b = (uint8_t*)dev_alloc(c, b_bytes + 16, &sz[1]);
    fsv_pair* d_pairs = (fsv_pair*)dev_alloc(c, n * sizeof(fsv_pair), &sz[2]);
    int32_t* d_order = (int32_t*)dev_alloc(c, n * 4, &sz[3]);
    int32_t* d_dist = (int32_t*)dev_alloc(c, n * 4, &sz[4]);
    int8_t* d_carry = (int8_t*)dev_alloc(c, (size_t)grid * warps_per_cta * (size_t)pitch, &sz[5]);
    unsigned int* d_cursor = (unsigned int*)dev_alloc(c, 64, &sz[6]);
    auto done = [&](int code) {
        dev_release(c, d_a, sz[0]); dev_release(c, d_b, sz[1]); dev_release(c, d_pairs, sz[2]); dev_release(c, d_order, sz[3]);
        dev_release(c, d_dist, sz[4]); dev_release(c, d_carry, sz[5]); dev_release(c, d_cursor, sz[6]);
        return code;
    };
    if (!d_a || !d_b || !d_pairs || !d_order || !d_dist || !d_carry || !d_cursor) return done(FSV_ERR_NOMEM);
#define CKE(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { c->last_error = std::string(#call) + " -> " + cudaGetErrorString(e_); cudaGetLastError(); return done(FSV_ERR_CUDA); } } while (0)
    CKE(cudaMemcpyAsync(d_a, ra.data(), a_bytes, cudaMemcpyHostToDevice, c->stream));
    CKE(cudaMemcpyAsync(d_b, rb.data(), b_bytes, cudaMemcpyHostToDevice, c->stream));
    CKE(cudaMemcpyAsync(d_pairs, pairs, n * sizeof(fsv_pair), cudaMemcpyHostToDevice, c->stream));
    CKE(cudaMemcpyAsync(d_order, order.data(), n * 4, cudaMemcpyHostToDevice, c->stream));
    CKE(cudaMemsetAsync(d_cursor, 0, 64, c->stream));
    EdParams P{d_a, d_b, d_pairs, d_order, (int32_t)n, d_dist, d_carry, pitch, d_cursor};
    cudaEvent_t e0, e1;
    CKE(cudaEventCreate(&e0)); CKE(cudaEventCreate(&e1));
    CKE(cudaEventRecord(e0, c->stream));
    fsv_edit_distance_kernel<<<grid, warps_per_cta * 32, 0, c->stream>>>(P);
    CKE(cudaGetLastError());
    CKE(cudaEventRecord(e1, c->stream));
    CKE(cudaMemcpyAsync(dist, d_dist, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CKE(cudaStreamSynchronize(c->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
#undef CKE
    c->stats.other_launches += 1;
    c->stats.total_ms = ms;
    c->stats.h2d_bytes += (int64_t)(a_bytes + b_bytes + n * (sizeof(fsv_pair) + 4));
    c->stats.d2h_bytes += (int64_t)n * 4;
    return done(FSV_OK);
}

extern "C" int fsv_batch_signatures(fsv_batch* b, const int64_t* ref_start, int min_svlen, fsv_signature* out, size_t cap, size_t* n_out)
{
    if (!b || !n_out || (cap && !out)) return FSV_ERR_INVALID;
    fsv_ctx* c = b->ctx;
    if (!c || b->state != 1) return FSV_ERR_STATE;
    *n_out = 0;
    const int n = (int)b->n;
    if (!n) return FSV_OK;
    CK(c, cudaSetDevice(c->device));
    size_t sz_counts = 0, sz_offs = 0, sz_ref = 0, sz_out = 0;
    int32_t* d_counts = (int32_t*)dev_alloc(c, (size_t)n * 8, &sz_counts);
    long long* d_offs = (long long*)dev_alloc(c, (size_t)n * 8, &sz_offs);
    long long* d_ref = ref_start ? (long long*)dev_alloc(c, (size_t)n * 8, &sz_ref) : nullptr;
    fsv_signature* d_out = nullptr;
    int rc = FSV_OK;
    auto done = [&](int code) {
        dev_release(c, d_counts, sz_counts); dev_release(c, d_offs, sz_offs); dev_release(c, d_ref, sz_ref); dev_release(c, d_out, sz_out);
        return code;
    };
    if (!d_counts || !d_offs || (ref_start && !d_ref)) return done(FSV_ERR_NOMEM);
#define CKS(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { c->last_error = std::string(#call) + " -> " + cudaGetErrorString(e_); cudaGetLastError(); return done(FSV_ERR_CUDA); } } while (0)
    if (ref_start) CKS(cudaMemcpyAsync(d_ref, ref_start, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    const int threads = 128, blocks = (n + threads - 1) / threads;
    fsv_sig_count_kernel<<<blocks, threads, 0, c->stream>>>(b->d_results, b->d_cigar, d_ref, n, min_svlen, d_counts);
    CKS(cudaGetLastError());
    std::vector<int32_t> counts((size_t)n * 2);
    CKS(cudaMemcpyAsync(counts.data(), d_counts, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CKS(cudaStreamSynchronize(c->stream));
    std::vector<long long> offs((size_t)n);
    long long total = 0;
    for (int i = 0; i < n; ++i) { offs[(size_t)i] = total; total += counts[2 * (size_t)i] + counts[2 * (size_t)i + 1]; }
    *n_out = (size_t)total;
    c->stats.other_launches += 1;
    c->stats.d2h_bytes += (int64_t)n * 8;
    if ((size_t)total > cap) return done(FSV_ERR_CIGAR_CAP);
    if (total) {
        d_out = (fsv_signature*)dev_alloc(c, (size_t)total * sizeof(fsv_signature), &sz_out);
        if (!d_out) return done(FSV_ERR_NOMEM);
        CKS(cudaMemcpyAsync(d_offs, offs.data(), (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
        fsv_sig_write_kernel<<<blocks, threads, 0, c->stream>>>(b->d_results, b->d_cigar, d_ref, n, min_svlen, d_counts, d_offs, d_out);
        CKS(cudaGetLastError());
        CKS(cudaMemcpyAsync(out, d_out, (size_t)total * sizeof(fsv_signature), cudaMemcpyDeviceToHost, c->stream));
        CKS(cudaStreamSynchronize(c->stream));
        c->stats.other_launches += 1;
        c->stats.d2h_bytes += total * (int64_t)sizeof(fsv_signature);
    }
#undef CKS
    return done(rc);
}

extern "C" int fsv_batch_timeline(fsv_batch* b, int64_t* start_end_ns)
{
    if (!b || !start_end_ns) return FSV_ERR_INVALID;
    fsv_ctx* c = b->ctx;
    if (!c || b->state != 1) return FSV_ERR_STATE;
    CK(c, cudaSetDevice(c->device));
    if (b->n) CK(c, cudaMemcpy(start_end_ns, b->d_timeline, b->n * 16, cudaMemcpyDeviceToHost));
    return FSV_OK;
}

extern "C" int fsv_batch_plan(const fsv_batch* b, int32_t* plan)
{
    if (!b || (!plan && b->n)) return FSV_ERR_INVALID;
    for (size_t i = 0; i < b->n; ++i) {
        int32_t v = 0;
        if (b->tasks[i].kind == 1) v |= b->is_dpx[i] ? FSV_PLAN_DPX : FSV_PLAN_GENERAL;
        if (b->tasks[i].kind == 1 && b->is_dpx[i] && b->tasks[i].pad_ > 0 && !(i < b->is_seg.size() && b->is_seg[i])) v |= FSV_PLAN_EDGE_WARP;
        if (i < b->is_seg.size() && b->is_seg[i]) v |= FSV_PLAN_SEGMENTED | ((int32_t)b->seg_tasks[(size_t)b->tasks[i].seg_id].n_segs << 16);
        if (i < b->is_excl.size() && b->is_excl[i]) v |= FSV_PLAN_EXCLUSIVE;
        v |= (b->tasks[i].nw & 0xf) << 8;
        plan[i] = v;
    }
    return FSV_OK;
}

extern "C" int fsv_align_batch(fsv_ctx* c, const fsv_scoring* sc,
                               const uint8_t* qarena, size_t qbytes, const uint8_t* tarena, size_t tbytes,
                               const fsv_task* tasks, size_t n, fsv_result* out,
                               uint32_t* cigar, size_t cigar_cap, size_t* cigar_used)
{
    fsv_batch* b = nullptr;
    int rc = fsv_batch_create(c, sc, qarena, qbytes, tarena, tbytes, tasks, n, &b);
    if (rc != FSV_OK) return rc;
    rc = fsv_batch_run(b);
    HostTrace tr;
    if (rc == FSV_OK) rc = fsv_batch_fetch(b, out, cigar, cigar_cap, cigar_used);
    tr.lap("fetch");
    fsv_batch_destroy(b);
    tr.lap("destroy");
    return rc;
}

static int single(fsv_ctx* c, const fsv_scoring& sc, int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                  int w, int zdrop, int end_bonus, int flag, fsv_result* ez, uint32_t* cigar, int cigar_cap)
{
    if (!c || !ez) return FSV_ERR_INVALID;
    fsv_task t{};
    t.q_off = 0; t.t_off = 0; t.qlen = qlen; t.tlen = tlen; t.w = w; t.zdrop = zdrop; t.end_bonus = end_bonus; t.flag = flag;
    size_t used = 0;
    return fsv_align_batch(c, &sc, query, qlen > 0 ? (size_t)qlen : 0, target, tlen > 0 ? (size_t)tlen : 0, &t, 1, ez,
                           cigar, cigar_cap > 0 ? (size_t)cigar_cap : 0, &used);
}

extern "C" int fsv_ksw_extz2(fsv_ctx* c, int qlen, const uint8_t* query, int tlen, const uint8_t* target, int8_t m,
                             const int8_t* mat, int8_t q, int8_t e, int w, int zdrop, int end_bonus, int flag,
                             fsv_result* ez, uint32_t* cigar, int cigar_cap)
{
    if (!mat || m > 5) return FSV_ERR_INVALID;
    fsv_scoring sc{};
    sc.m = m; sc.q = q; sc.e = e; sc.q2 = -1; sc.e2 = -1;
    if (m > 0) memcpy(sc.mat, mat, (size_t)m * m);
    return single(c, sc, qlen, query, tlen, target, w, zdrop, end_bonus, flag, ez, cigar, cigar_cap);
}

extern "C" int fsv_ksw_extd2(fsv_ctx* c, int qlen, const uint8_t* query, int tlen, const uint8_t* target, int8_t m,
                             const int8_t* mat, int8_t q, int8_t e, int8_t q2, int8_t e2, int w, int zdrop,
                             int end_bonus, int flag, fsv_result* ez, uint32_t* cigar, int cigar_cap)
{
    if (!mat || m > 5 || q2 < 0) return FSV_ERR_INVALID;
    fsv_scoring sc{};
    sc.m = m; sc.q = q; sc.e = e; sc.q2 = q2; sc.e2 = e2;
    if (m > 0) memcpy(sc.mat, mat, (size_t)m * m);
    return single(c, sc, qlen, query, tlen, target, w, zdrop, end_bonus, flag, ez, cigar, cigar_cap);
}

// ---------------------------------------------------------------------------
// Level 1: presets and the region hook (host composition over fsv_align_batch)
static const fsv_preset kPresets[] = {
    // name        a  b   q   e  q2  e2  zdrop zinv  bw    bw_long sc_ambi end_bonus      (minimap2 2.24 options.c values)
    {"asm5",       1, 19, 39, 3, 81, 1,  200,  200,  2000, 100000, 1, -1},
    {"asm10",      1, 9,  16, 2, 41, 1,  200,  200,  2000, 100000, 1, -1},
    {"map-hifi",   1, 4,  6,  2, 26, 1,  400,  200,  2000, 20000,  1, -1},
    {"map-pb",     2, 4,  4,  2, 24, 1,  400,  200,  2000, 20000,  1, -1},
    {"map-ont",    2, 4,  4,  2, 24, 1,  400,  200,  2000, 20000,  1, -1},
    {"hifiasm",    2, 4,  4,  2, -1, -1, 400,  400,  500,  500,    0, 0},      // Correct.h:1194-1199, Correct.cpp:7670
};

extern "C" int fsv_preset_lookup(const char* name, fsv_preset* out, fsv_scoring* sc)
{
    if (!name) return FSV_ERR_INVALID;
    for (const fsv_preset& p : kPresets) {
        if (strcmp(p.name, name)) continue;
        if (out) *out = p;
        if (sc) {
            memset(sc, 0, sizeof *sc);
            sc->m = 5; sc->q = (int8_t)p.q; sc->e = (int8_t)p.e; sc->q2 = (int8_t)p.q2; sc->e2 = (int8_t)p.e2;
            for (int i = 0; i < 5; ++i)
                for (int j = 0; j < 5; ++j)
                    sc->mat[i * 5 + j] = (int8_t)((i == 4 || j == 4) ? -abs(p.sc_ambi) : i == j ? abs(p.a) : -abs(p.b));
        }
        return FSV_OK;
    }
    return FSV_ERR_INVALID;
}

extern "C" int fsv_realign_regions(fsv_ctx* c, const uint8_t* ref_codes, size_t ref_len,
                                   const int64_t* region_start, const int64_t* region_end,
                                   const uint8_t* contig_codes, size_t contig_bytes,
                                   const int64_t* contig_off, const int32_t* contig_len, size_t n,
                                   const char* preset, int bw, int flag,
                                   fsv_record* out, uint32_t* cigar, size_t cigar_cap, size_t* cigar_used)
{
    if (!c) return FSV_ERR_INVALID;
    if (cigar_used) *cigar_used = 0;
    fsv_preset p; fsv_scoring sc;
    if (fsv_preset_lookup(preset, &p, &sc) != FSV_OK) { c->last_error = "unknown preset"; return FSV_ERR_INVALID; }
    if (n == 0) return FSV_OK;
    if (!ref_codes || !region_start || !region_end || !contig_codes || !contig_off || !contig_len || !out || bw < 0)
        return FSV_ERR_INVALID;
    const int w = (int)(bw * 1.5 + 1.0);                    // minimap2 runs ksw2 with bw*1.5+1
    std::vector<fsv_task> tasks(n);
    for (size_t i = 0; i < n; ++i) {
        const int64_t s = region_start[i], e = region_end[i];
        if (s < 0 || e < s || (size_t)e > ref_len || e - s > INT32_MAX || contig_len[i] < 0 || contig_off[i] < 0 ||
            (size_t)(contig_off[i] + contig_len[i]) > contig_bytes) {
            c->last_error = "fsv_realign_regions: region or contig " + std::to_string(i) + " outside its arena";
            return FSV_ERR_INVALID;
        }
        fsv_task& t = tasks[i];
        t.q_off = contig_off[i]; t.t_off = s; t.qlen = contig_len[i]; t.tlen = (int32_t)(e - s);
        t.w = w; t.zdrop = p.zdrop; t.end_bonus = 0; t.flag = flag & ~FSV_EZ_SCORE_ONLY;
    }
    std::vector<fsv_result> res(n);
    const int rc = fsv_align_batch(c, &sc, contig_codes, contig_bytes, ref_codes, ref_len, tasks.data(), n, res.data(),
                                   cigar, cigar_cap, cigar_used);
    if (rc != FSV_OK && rc != FSV_ERR_CIGAR_CAP) return rc;
    for (size_t i = 0; i < n; ++i) {
        fsv_record& r = out[i];
        r.pos = region_start[i]; r.ref_end = region_start[i];
        r.cigar_off = res[i].cigar_off; r.n_cigar = res[i].n_cigar; r.query_length = contig_len[i];
        r.score = res[i].score; r.zdropped = res[i].zdropped; r.is_reverse = 0; r.mapq = res[i].zdropped ? 0 : 60;      // a dropped task stops at its maximum cell: not a full-length alignment
        if (rc == FSV_OK)
            for (int32_t k = 0; k < r.n_cigar; ++k) {
                const uint32_t wd = cigar[r.cigar_off + k];
                if ((wd & 0xf) == 0 || (wd & 0xf) == 2) r.ref_end += wd >> 4;
            }
    }
    return rc;
}

// ---------------------------------------------------------------------------
// roofline denominators measured on this device (see fsv_peaks.cuh)
template <int KIND>
static int run_peak(fsv_ctx* c, int iters, double* out)
{
    uint32_t* d = nullptr;
    CK(c, cudaMalloc(&d, 4096));
    int grid = c->sm_count * 8;
    cudaEvent_t e0, e1;
    CK(c, cudaEventCreate(&e0)); CK(c, cudaEventCreate(&e1));
    fsv_peak_kernel<KIND><<<grid, 256, 0, c->stream>>>(d, iters / 8, 12345u);      // warm-up
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        CK(c, cudaEventRecord(e0, c->stream));
        fsv_peak_kernel<KIND><<<grid, 256, 0, c->stream>>>(d, iters, 12345u + rep);
        CK(c, cudaEventRecord(e1, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
        float ms = 0;
        CK(c, cudaEventElapsedTime(&ms, e0, e1));
        double ops = (double)grid * 256.0 * (double)iters * PEAK_OPS_PER_ITER;
        best = std::max(best, ops / (ms * 1e-3));
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *out = best;
    return FSV_OK;
}

extern "C" int fsv_measure_int_peak(fsv_ctx* c, int kind, double* lane_ops_per_s)
{
    if (!c || !lane_ops_per_s) return FSV_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    const int iters = 4096;
    switch (kind) {
        case 0: return run_peak<0>(c, iters, lane_ops_per_s);
        case 1: return run_peak<1>(c, iters, lane_ops_per_s);
        case 2: return run_peak<2>(c, iters, lane_ops_per_s);
        case 3: return run_peak<3>(c, iters, lane_ops_per_s);
        case 4: return run_peak<4>(c, iters, lane_ops_per_s);
        case 5: return run_peak<5>(c, iters, lane_ops_per_s);
        case 6: return run_peak<6>(c, iters, lane_ops_per_s);
        case 7: return run_peak<7>(c, iters, lane_ops_per_s);
        default: return FSV_ERR_INVALID;
    }
}
