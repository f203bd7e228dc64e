// fsv_common.cuh — shared device/host declarations of libfocalsv_cuda (sm_100a).
//
// Vocabulary follows the reference (software/hifiasm-0.16.1/ksw2_extz2_sse.c):
//   r      antidiagonal index, 0 .. qlen+tlen-2              (:101)
//   t      target column; query row j = r - t
//   st0/en0  exact band limits of antidiagonal r             (:102-115)
//   st/en    the same rounded to 16-lane vectors             (:116)
//   u v x y (x2 y2)  the Suzuki-Kasahara difference arrays   (:26-47)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/focalsv_cuda.h"

namespace fsv {

// one task as the kernels see it (built on the host from fsv_task)
struct DevTask {
    int64_t q_off, t_off;   // byte offsets into the device sequence arenas
    int64_t tb_off;         // byte offset of this task's traceback rows, -1 = score only
    int64_t cells_est;      // in-band cells of a full run (sort key)
    int32_t qlen, tlen;
    int32_t w;              // effective band (w < 0 already replaced, ksw2_extz2_sse.c:72)
    int32_t zdrop, end_bonus, flag;
    int32_t pitch;          // traceback bytes per antidiagonal = n_col_*16 (:75-76)
    int32_t orig;           // index in the caller's task array
    int32_t kind;           // 0 = reset result only (ksw2's silent returns), 1 = run
    int32_t pad_;           // kind 0: the status to report
    int32_t tb_mode;        // traceback direction encoding: 0 = ksw2's d, n > 0 = (n - d) (DPX kernel)
    int32_t nw;             // DPX kernel: warps per task (0 = general kernel)
    int32_t tb_pages;       // traceback pages this task needs (0 = score only)
    int32_t rows_per_page;  // antidiagonals per page
};

// where the CIGAR walk of a task starts (ksw2_extz2_sse.c:292-301)
struct DevAux {
    int32_t i0, j0;         // start cell (target, query); i0 < 0 = no CIGAR
    int32_t n_cigar;        // filled by the counting pass
    int32_t pad_;
};

struct DevScoring {
    int32_t m, dual;
    int32_t q, e, q2, e2;           // dual: already swapped so that q+e <= q2+e2
    int32_t sc_mch, sc_mis, sc_N;   // as int8 bit patterns widened to int
    int32_t max_sc_clamp;           // extz2: (int8)(mat[0] + 2(q+e)) ; extd2: mat[0]
    int32_t long_thres, long_diff;  // extd2 first row/column envelope
    int32_t e_drop;                 // gap slack of the z-drop test (e, or e2 for dual)
    int8_t mat[32];
};

// band limits of antidiagonal r (ksw2_extz2_sse.c:102-110)
__host__ __device__ __forceinline__ void band_limits(int r, int qlen, int tlen, int w, int& st0, int& en0)
{
    int s = 0, e = tlen - 1;
    if (s < r - qlen + 1) s = r - qlen + 1;
    if (e > r) e = r;
    int lo = (r - w + 1) >> 1, hi = (r + w) >> 1;
    if (s < lo) s = lo;
    if (e > hi) e = hi;
    st0 = s; en0 = e;
}

__host__ __device__ __forceinline__ int round_st(int st0) { return st0 / 16 * 16; }
__host__ __device__ __forceinline__ int round_en(int en0) { return (en0 + 16) / 16 * 16 - 1; }

// ksw_extz_t bookkeeping kept in registers by every thread of a task's CTA
struct EzState {
    int32_t max, max_t, max_q, mqe, mqe_t, mte, mte_q, score, zdropped;
    __device__ __forceinline__ void reset()
    {   // ksw2.h:153-158
        max_q = max_t = mqe_t = mte_q = -1;
        max = 0; score = mqe = mte = FSV_NEG_INF; zdropped = 0;
    }
    // ksw2.h:160-176 (is_rot = 1); returns 1 when the extension is dropped
    __device__ __forceinline__ int apply_zdrop(int32_t H, int r, int t, int zdrop, int e)
    {
        if (H > max) {
            max = (int32_t)((uint32_t)H & 0x7fffffffu); max_t = t; max_q = r - t;
        } else if (t >= max_t && r - t >= max_q) {
            int tl = t - max_t, ql = (r - t) - max_q;
            int l = tl > ql ? tl - ql : ql - tl;
            if (zdrop >= 0 && max - H > zdrop + l * e) { zdropped = 1; return 1; }
        }
        return 0;
    }
};

}  // namespace fsv

namespace fsv {

// ---------------------------------------------------------------------------
// Paged traceback pool.  Traceback rows (one per antidiagonal, `pitch` bytes) of a task live in
// fixed-size pages taken from one device-wide pool when the task STARTS and given back when its
// CIGAR has been reconstructed, so the resident traceback is bounded by the tasks in flight
// (one per CTA), not by the batch.  Row r of a task sits in page r / rows_per_page.
struct TbPool {
    uint8_t* base;          // n_pages * page_bytes
    int64_t page_bytes;
    int32_t n_pages;
    int32_t* free_stack;    // page ids, free_stack[0 .. *n_free)
    int32_t* n_free;
    int32_t* lock;          // 0 = free
};

__device__ __forceinline__ void pool_lock(int32_t* lock)
{
    unsigned ns = 64;
    while (atomicCAS(lock, 0, 1) != 0) { __nanosleep(ns); if (ns < 4096) ns <<= 1; }
    __threadfence();
}
__device__ __forceinline__ void pool_unlock(int32_t* lock)
{
    __threadfence();
    atomicExch(lock, 0);
}
// called by ONE thread; returns false when fewer than n pages are free
__device__ inline bool pool_try_alloc(const TbPool& P, int n, int32_t* table)
{
    if (n <= 0) return true;
    bool ok = false;
    pool_lock(P.lock);
    int nf = *(volatile int32_t*)P.n_free;
    if (nf >= n) {
        for (int i = 0; i < n; ++i) table[i] = ((volatile int32_t*)P.free_stack)[nf - 1 - i];
        *(volatile int32_t*)P.n_free = nf - n;
        ok = true;
    }
    pool_unlock(P.lock);
    return ok;
}
__device__ inline void pool_free(const TbPool& P, int n, const int32_t* table)
{
    if (n <= 0) return;
    pool_lock(P.lock);
    int nf = *(volatile int32_t*)P.n_free;
    for (int i = 0; i < n; ++i) ((volatile int32_t*)P.free_stack)[nf + i] = table[i];
    *(volatile int32_t*)P.n_free = nf + n;
    pool_unlock(P.lock);
}

// ---------------------------------------------------------------------------
// Two-ended work queue over a task list sorted largest-first: CTAs take from the head; a CTA whose
// head task does not get its traceback pages yet keeps it pending and works on tasks from the TAIL
// (the smallest ones) meanwhile, so the few very long tasks never leave the device idle.
struct TaskQueue {
    unsigned long long* state;   // (head << 32) | tail over order[0..n)
    const int32_t* order;
};
// end: 0 = head, 1 = tail.  Returns the task index or -1 when the queue is empty.
__device__ inline int queue_take(const TaskQueue& Q, int end)
{
    unsigned long long s = *(volatile unsigned long long*)Q.state;
    for (;;) {
        unsigned head = (unsigned)(s >> 32), tail = (unsigned)(s & 0xffffffffu);
        if (head >= tail) return -1;
        unsigned long long ns = end == 0 ? (((unsigned long long)(head + 1) << 32) | tail)
                                         : (((unsigned long long)head << 32) | (tail - 1));
        unsigned long long old = atomicCAS(Q.state, s, ns);
        if (old == s) return Q.order[end == 0 ? head : tail - 1];
        s = old;
    }
}

// everything a fill kernel needs besides its own parameters
struct RunCtx {
    const uint8_t* qarena;
    const uint8_t* tarena;
    const DevTask* tasks;
    fsv_result* results;
    TbPool pool;
    int32_t* page_tables;        // per CTA: max_pages_per_task entries
    int32_t max_pages_per_task;
    uint32_t* cigar;             // compact CIGAR arena
    unsigned long long* cigar_cursor;
    int64_t cigar_cap;           // words
    int32_t* overflow;           // set when the arena is too small
    long long* timeline;         // per task (caller order): start ns, end ns (globaltimer), or null
    DevScoring sc;
};

// Picks the next task for this CTA (thread 0 only) and gets its traceback pages.
// `pending` carries a claimed task that is still waiting for memory.  Returns the task index or -1.
__device__ inline int next_task(const RunCtx& C, const TaskQueue& Q, int32_t* table, int& pending)
{
    for (;;) {
        if (pending >= 0) {
            if (pool_try_alloc(C.pool, C.tasks[pending].tb_pages, table)) { int t = pending; pending = -1; return t; }
            int t = queue_take(Q, 1);                       // memory is short: do a small task meanwhile
            if (t >= 0) {
                if (pool_try_alloc(C.pool, C.tasks[t].tb_pages, table)) return t;
                // not even the small one fits right now: wait for running tasks to finish
                for (;;) { __nanosleep(2000); if (pool_try_alloc(C.pool, C.tasks[t].tb_pages, table)) return t; }
            }
            __nanosleep(2000);                              // queue drained: wait for memory
            continue;
        }
        int t = queue_take(Q, 0);
        if (t < 0) return -1;
        if (pool_try_alloc(C.pool, C.tasks[t].tb_pages, table)) return t;
        pending = t;
    }
}

__device__ __forceinline__ long long global_ns()
{
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// byte address of traceback row r of a task (rows_per_page = page_bytes / pitch)
__device__ __forceinline__ uint8_t* tb_row(const TbPool& P, const int32_t* table, int rows_per_page, int pitch, int r)
{
    const int pg = r / rows_per_page;
    return P.base + (int64_t)table[pg] * P.page_bytes + (int64_t)(r - pg * rows_per_page) * pitch;
}

}  // namespace fsv
