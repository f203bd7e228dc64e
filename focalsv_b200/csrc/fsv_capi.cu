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
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "fsv_backtrack.cuh"
#include "fsv_common.cuh"
#include "fsv_fill_dpx.cuh"
#include "fsv_fill_exact.cuh"
#include "fsv_peaks.cuh"

using namespace fsv;

struct fsv_ctx {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    std::string last_error;
    fsv_stats stats{};
    // options
    int64_t tb_budget = 0;      // bytes of traceback kept resident per chunk (0 = auto)
    int force_exact = 0;        // route every task to the general int8-exact kernel
    int exact_smem_lanes = 4096;
    // scratch shared by the batches of this context (one batch runs at a time)
    uint8_t* d_tb = nullptr; size_t tb_cap = 0;
    uint8_t* d_ws = nullptr; size_t ws_cap = 0;
};

struct Chunk { int32_t begin, end; int64_t tb_bytes; };
struct DpxLaunch { int chunk, nw, with_tb, begin, count; };   // a slice of order_dpx for one kernel variant

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
    std::vector<int32_t> order_exact; // per chunk: tasks for the general kernel
    std::vector<int32_t> order_all;   // processing order (largest first), chunked
    std::vector<int32_t> order_dpx;   // per chunk: tasks for the DPX kernel (same index space as order_all)
    std::vector<Chunk> chunks;
    std::vector<DpxLaunch> dpx_launches;
    std::vector<uint8_t> is_dpx;      // per task
    int64_t cigar_cap_words = 0;
    int64_t ws_lanes = 0;
    // device
    uint8_t *d_q = nullptr, *d_t = nullptr;
    DevTask* d_tasks = nullptr;
    int32_t *d_order_all = nullptr, *d_order_exact = nullptr, *d_order_dpx = nullptr;
    fsv_result* d_results = nullptr;
    DevAux* d_aux = nullptr;
    int32_t* d_counters = nullptr;    // [0] exact cursor, [1] dpx cursor, [2] overflow flag
    int64_t* d_running = nullptr;
    uint32_t* d_cigar = nullptr;
    int state = 0;                    // 0 created, 1 run
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
    *out = c;
    return FSV_OK;
}

extern "C" void fsv_destroy(fsv_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->d_tb) cudaFree(c->d_tb);
    if (c->d_ws) cudaFree(c->d_ws);
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
    cudaFree(b->d_q); cudaFree(b->d_t); cudaFree(b->d_tasks); cudaFree(b->d_order_all); cudaFree(b->d_order_exact);
    cudaFree(b->d_order_dpx); cudaFree(b->d_results); cudaFree(b->d_aux); cudaFree(b->d_counters);
    cudaFree(b->d_running); cudaFree(b->d_cigar);
}

extern "C" void fsv_batch_destroy(fsv_batch* b)
{
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    free_batch_device(b);
    delete b;
}

extern "C" int fsv_batch_create(fsv_ctx* c, const fsv_scoring* scoring,
                                const uint8_t* qarena, size_t qbytes, const uint8_t* tarena, size_t tbytes,
                                const fsv_task* tasks, size_t n, fsv_batch** out)
{
    if (!c || !scoring || !out || (n && !tasks)) return FSV_ERR_INVALID;
    *out = nullptr;
    if (n > 0x7ffffff0u) return FSV_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    fsv_batch* b = new (std::nothrow) fsv_batch();
    if (!b) return FSV_ERR_NOMEM;
    b->ctx = c; b->n = n;
    bool all_reset; int reset_status;
    int rc = build_scoring(scoring, &b->sc, &b->dual, &reset_status, &all_reset);
    if (rc != FSV_OK) { delete b; return rc; }

    // ---- task table
    b->tasks.resize(n);
    b->is_dpx.assign(n, 0);
    int64_t cigar_words = 0;
    for (size_t i = 0; i < n; ++i) {
        const fsv_task& t = tasks[i];
        DevTask& d = b->tasks[i];
        memset(&d, 0, sizeof d);
        d.q_off = t.q_off; d.t_off = t.t_off; d.qlen = t.qlen; d.tlen = t.tlen;
        d.zdrop = t.zdrop; d.end_bonus = t.end_bonus; d.flag = t.flag; d.orig = (int32_t)i; d.tb_off = -1;
        d.kind = 1;
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
        if (!(t.flag & FSV_EZ_SCORE_ONLY)) cigar_words += (int64_t)t.qlen + t.tlen + 2;
        if (!c->force_exact) {
            // the DPX kernel packs bases in 2 bits: a task with a wildcard base goes to the general kernel
            const bool wild = has_wildcard(qarena + t.q_off, (size_t)t.qlen) || has_wildcard(tarena + t.t_off, (size_t)t.tlen);
            if (dpx_supports(b->sc, d, wild)) {
                b->is_dpx[i] = 1;
                d.nw = dpx_class_of(dpx_warps_needed(d));
                d.tb_mode = b->dual ? 4 : 2;
            }
        }
    }
    b->cigar_cap_words = cigar_words;

    // ---- processing order: largest first (LPT within the device), cut into chunks
    // whose traceback rows fit the resident budget
    std::vector<int32_t> ord(n);
    for (size_t i = 0; i < n; ++i) ord[i] = (int32_t)i;
    std::stable_sort(ord.begin(), ord.end(), [&](int32_t a, int32_t x) { return b->tasks[a].cells_est > b->tasks[x].cells_est; });
    int64_t budget = c->tb_budget;
    if (budget <= 0) {
        size_t fr = 0, tot = 0;
        CK(c, cudaMemGetInfo(&fr, &tot));
        budget = (int64_t)(fr * 0.70);
    }
    b->order_all = ord;
    {
        Chunk ch{0, 0, 0};
        for (size_t k = 0; k < n; ++k) {
            DevTask& d = b->tasks[ord[k]];
            int64_t need = 0;
            if (d.kind == 1 && !(d.flag & FSV_EZ_SCORE_ONLY))
                need = ((int64_t)(d.qlen + d.tlen - 1) * d.pitch + 255) / 256 * 256;
            if (need > budget) {
                c->last_error = "traceback of task " + std::to_string(d.orig) + " (" + std::to_string(need) +
                                " bytes) exceeds the resident budget";
                delete b; return FSV_ERR_NOMEM;
            }
            if (ch.tb_bytes + need > budget && (int32_t)k > ch.begin) {
                ch.end = (int32_t)k; b->chunks.push_back(ch);
                ch = Chunk{(int32_t)k, 0, 0};
            }
            if (need) { d.tb_off = ch.tb_bytes; ch.tb_bytes += need; }
        }
        ch.end = (int32_t)n;
        if (ch.end > ch.begin) b->chunks.push_back(ch);
    }
    // per chunk, split the order into work lists: one per kernel variant (general; DPX by
    // warps-per-task class and with/without traceback), each still largest-first
    b->order_exact.assign(n, -1); b->order_dpx.assign(n, -1);
    int64_t ws_need = 0;
    for (size_t ci = 0; ci < b->chunks.size(); ++ci) {
        auto& ch = b->chunks[ci];
        int ne = 0, nd = 0;
        for (int k = ch.begin; k < ch.end; ++k) {
            int ti = ord[k];
            if (b->is_dpx[ti]) continue;
            b->order_exact[ch.begin + ne++] = ti;
            if (b->tasks[ti].kind == 1 && b->tasks[ti].pitch + 96 > c->exact_smem_lanes)
                ws_need = std::max<int64_t>(ws_need, b->tasks[ti].pitch + 96);
        }
        static const int kClasses[5] = {8, 6, 4, 2, 1};
        for (int cls : kClasses)
            for (int with_tb = 1; with_tb >= 0; --with_tb) {
                int begin = ch.begin + nd;
                for (int k = ch.begin; k < ch.end; ++k) {
                    int ti = ord[k];
                    if (!b->is_dpx[ti] || b->tasks[ti].nw != cls) continue;
                    if ((int)!(b->tasks[ti].flag & FSV_EZ_SCORE_ONLY) != with_tb) continue;
                    b->order_dpx[ch.begin + nd++] = ti;
                }
                if (ch.begin + nd > begin) b->dpx_launches.push_back(DpxLaunch{(int)ci, cls, with_tb, begin, ch.begin + nd - begin});
            }
    }
    b->ws_lanes = ws_need ? pow2_at_least(ws_need) : 0;

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
    CKB(cudaMalloc(&b->d_q, qbytes + 64));
    CKB(cudaMalloc(&b->d_t, tbytes + 64));
    CKB(cudaMalloc(&b->d_tasks, (n + 1) * sizeof(DevTask)));
    CKB(cudaMalloc(&b->d_order_all, (n + 1) * 4));
    CKB(cudaMalloc(&b->d_order_exact, (n + 1) * 4));
    CKB(cudaMalloc(&b->d_order_dpx, (n + 1) * 4));
    CKB(cudaMalloc(&b->d_results, (n + 1) * sizeof(fsv_result)));
    CKB(cudaMalloc(&b->d_aux, (n + 1) * sizeof(DevAux)));
    CKB(cudaMalloc(&b->d_counters, 64));
    CKB(cudaMalloc(&b->d_running, 64));
    CKB(cudaMalloc(&b->d_cigar, (size_t)(b->cigar_cap_words + 4) * 4));
    if (qbytes) CKB(cudaMemcpyAsync(b->d_q, qarena, qbytes, cudaMemcpyHostToDevice, c->stream));
    if (tbytes) CKB(cudaMemcpyAsync(b->d_t, tarena, tbytes, cudaMemcpyHostToDevice, c->stream));
    if (n) {
        CKB(cudaMemcpyAsync(b->d_tasks, b->tasks.data(), n * sizeof(DevTask), cudaMemcpyHostToDevice, c->stream));
        CKB(cudaMemcpyAsync(b->d_order_all, b->order_all.data(), n * 4, cudaMemcpyHostToDevice, c->stream));
        CKB(cudaMemcpyAsync(b->d_order_exact, b->order_exact.data(), n * 4, cudaMemcpyHostToDevice, c->stream));
        CKB(cudaMemcpyAsync(b->d_order_dpx, b->order_dpx.data(), n * 4, cudaMemcpyHostToDevice, c->stream));
    }
    CKB(cudaStreamSynchronize(c->stream));
#undef CKB
    c->stats.h2d_bytes += (int64_t)(qbytes + tbytes + n * (sizeof(DevTask) + 12));
    *out = b;
    return FSV_OK;
}

static int ensure_scratch(fsv_ctx* c, size_t tb_bytes, size_t ws_bytes)
{
    if (tb_bytes > c->tb_cap) {
        if (c->d_tb) { cudaFree(c->d_tb); c->d_tb = nullptr; c->tb_cap = 0; }
        CK(c, cudaMalloc(&c->d_tb, tb_bytes + 256));
        c->tb_cap = tb_bytes;
    }
    if (ws_bytes > c->ws_cap) {
        if (c->d_ws) { cudaFree(c->d_ws); c->d_ws = nullptr; c->ws_cap = 0; }
        CK(c, cudaMalloc(&c->d_ws, ws_bytes + 256));
        c->ws_cap = ws_bytes;
    }
    return FSV_OK;
}

template <bool DUAL>
static int launch_exact(fsv_ctx* c, fsv_batch* b, const Chunk& ch, int n_exact)
{
    if (n_exact <= 0) return FSV_OK;
    const size_t smem = (size_t)c->exact_smem_lanes * (EXACT_NARR + 4);
    auto kern = fsv_fill_exact_kernel<DUAL>;
    CK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, EXACT_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    int grid = std::min(n_exact, c->sm_count * per_sm);
    size_t ws_bytes = b->ws_lanes ? (size_t)grid * (size_t)b->ws_lanes * (EXACT_NARR + 4) : 0;
    int rc = ensure_scratch(c, 0, ws_bytes);
    if (rc != FSV_OK) return rc;
    FillParams P{};
    P.qarena = b->d_q; P.tarena = b->d_t; P.tasks = b->d_tasks;
    P.order = b->d_order_exact + ch.begin; P.n_order = n_exact; P.counter = b->d_counters + 0;
    P.results = b->d_results; P.aux = b->d_aux; P.tb = c->d_tb; P.ws = c->d_ws; P.ws_lanes = b->ws_lanes;
    P.smem_lanes = c->exact_smem_lanes; P.sc = b->sc;
    kern<<<grid, EXACT_THREADS, smem, c->stream>>>(P);
    CK(c, cudaGetLastError());
    c->stats.fill_launches++;
    return FSV_OK;
}

extern "C" int fsv_batch_run(fsv_batch* b)
{
    if (!b) return FSV_ERR_INVALID;
    fsv_ctx* c = b->ctx;
    CK(c, cudaSetDevice(c->device));
    cudaEvent_t e0, e1;
    CK(c, cudaEventCreate(&e0)); CK(c, cudaEventCreate(&e1));
    std::vector<cudaEvent_t> evs;   // per chunk: fill begin, fill end, backtrack end
    size_t tb_max = 0;
    for (auto& ch : b->chunks) tb_max = std::max<size_t>(tb_max, (size_t)ch.tb_bytes);
    int rc = ensure_scratch(c, tb_max, 0);
    if (rc != FSV_OK) return rc;
    CK(c, cudaMemsetAsync(b->d_running, 0, 64, c->stream));
    CK(c, cudaMemsetAsync(b->d_counters, 0, 64, c->stream));
    CK(c, cudaEventRecord(e0, c->stream));
    int64_t tb_total = 0;
    for (size_t ci = 0; ci < b->chunks.size(); ++ci) {
        auto& ch = b->chunks[ci];
        int n_exact = 0;
        for (int k = ch.begin; k < ch.end; ++k) if (b->order_exact[k] >= 0) ++n_exact;
        cudaEvent_t a, m, z;
        CK(c, cudaEventCreate(&a)); CK(c, cudaEventCreate(&m)); CK(c, cudaEventCreate(&z));
        evs.push_back(a); evs.push_back(m); evs.push_back(z);
        CK(c, cudaMemsetAsync(b->d_counters, 0, 8, c->stream));          // [0] general-kernel cursor
        CK(c, cudaMemsetAsync(b->d_counters + 4, 0, 48, c->stream));     // [4..15] one cursor per DPX launch
        CK(c, cudaEventRecord(a, c->stream));
        int slot = 4;
        for (const DpxLaunch& L : b->dpx_launches) {
            if (L.chunk != (int)ci) continue;
            DpxParams D{};
            D.qarena = b->d_q; D.tarena = b->d_t; D.tasks = b->d_tasks; D.order = b->d_order_dpx + L.begin;
            D.n_order = L.count; D.counter = b->d_counters + slot++; D.results = b->d_results; D.aux = b->d_aux;
            D.tb = c->d_tb; D.sc = b->sc;
            rc = dpx_launch(c->stream, c->sm_count, b->dual, L.with_tb != 0, L.nw, D, &c->last_error);
            if (rc != FSV_OK) return rc;
            c->stats.fill_launches++;
        }
        rc = b->dual ? launch_exact<true>(c, b, ch, n_exact) : launch_exact<false>(c, b, ch, n_exact);
        if (rc != FSV_OK) return rc;
        CK(c, cudaEventRecord(m, c->stream));
        // CIGAR reconstruction: count -> offsets -> write
        int n_all = ch.end - ch.begin;
        BtParams Q{};
        Q.tasks = b->d_tasks; Q.order = b->d_order_all + ch.begin; Q.n_order = n_all; Q.results = b->d_results;
        Q.aux = b->d_aux; Q.tb = c->d_tb; Q.cigar = b->d_cigar; Q.cigar_cap = b->cigar_cap_words;
        Q.overflow = b->d_counters + 2;
        int bt_grid = (n_all * 32 + BT_THREADS - 1) / BT_THREADS;
        fsv_backtrack_kernel<false><<<bt_grid, BT_THREADS, 0, c->stream>>>(Q);
        fsv_cigar_offsets_kernel<<<1, 1024, 0, c->stream>>>(b->d_tasks, b->d_order_all + ch.begin, n_all, b->d_aux,
                                                            b->d_results, b->d_running);
        fsv_backtrack_kernel<true><<<bt_grid, BT_THREADS, 0, c->stream>>>(Q);
        CK(c, cudaGetLastError());
        c->stats.backtrack_launches += 2; c->stats.other_launches += 1;
        CK(c, cudaEventRecord(z, c->stream));
        tb_total += ch.tb_bytes;
    }
    CK(c, cudaEventRecord(e1, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    float ms = 0;
    CK(c, cudaEventElapsedTime(&ms, e0, e1));
    c->stats.total_ms = ms; c->stats.fill_ms = 0; c->stats.backtrack_ms = 0;
    for (size_t i = 0; i + 2 < evs.size() + 0; i += 3) {
        float f = 0, g = 0;
        cudaEventElapsedTime(&f, evs[i], evs[i + 1]); cudaEventElapsedTime(&g, evs[i + 1], evs[i + 2]);
        c->stats.fill_ms += f; c->stats.backtrack_ms += g;
    }
    for (auto e : evs) cudaEventDestroy(e);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    c->stats.traceback_bytes = tb_total;
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
    if (b->state != 1) return FSV_ERR_STATE;
    CK(c, cudaSetDevice(c->device));
    int64_t used = 0;
    if (b->n) CK(c, cudaMemcpyAsync(out, b->d_results, b->n * sizeof(fsv_result), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpyAsync(&used, b->d_running, 8, cudaMemcpyDeviceToHost, c->stream));
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

extern "C" int fsv_align_batch(fsv_ctx* c, const fsv_scoring* sc,
                               const uint8_t* qarena, size_t qbytes, const uint8_t* tarena, size_t tbytes,
                               const fsv_task* tasks, size_t n, fsv_result* out,
                               uint32_t* cigar, size_t cigar_cap, size_t* cigar_used)
{
    fsv_batch* b = nullptr;
    int rc = fsv_batch_create(c, sc, qarena, qbytes, tarena, tbytes, tasks, n, &b);
    if (rc != FSV_OK) return rc;
    rc = fsv_batch_run(b);
    if (rc == FSV_OK) rc = fsv_batch_fetch(b, out, cigar, cigar_cap, cigar_used);
    fsv_batch_destroy(b);
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
