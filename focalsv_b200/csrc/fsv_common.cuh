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
    int32_t wild;           // some base of the task is the wildcard (code > 3)
    int32_t seg_id;         // index into RunCtx::seg_tasks when the task is cut into segments, else -1
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
// fixed-size pages of one device-wide pool and are given back when the task's CIGAR has been
// reconstructed, so the resident traceback is bounded by the tasks in flight (one per CTA), not by the
// batch.  Row r of a task sits in page r / rows_per_page.
//   * short tasks take all their pages when they START ("up front");
//   * long tasks (>= lazy_min_pages) take theirs ONE BY ONE as their antidiagonals advance ("lazy"): a
//     task that runs for seconds then holds on average half of its traceback, which is what bounds how
//     many long tasks can overlap.  Lazy growth could deadlock (every running task waiting for a page), so
//     every grant — a lazy page, the admission of a lazy task, the pages of a short task — is made only
//     if the lazy tasks can still all finish one after the other with the pages that remain free
//     (banker's algorithm, exact: the list of lazy tasks is short and scanned under the pool lock).
//     Admission of a lazy task additionally projects the growth of the lazy tasks already running
//     (same antidiagonal rate for all) and waits while the projected peak exceeds `lazy_fill` of the pool,
//     which staggers the long tasks instead of letting them hit the limit together.
constexpr int LAZY_MAX = 256;        // lazy tasks in flight (more wait for a slot)
constexpr int LAZY_BUCKETS = 32;     // time buckets of the growth projection
struct LazyState {                   // device global memory, touched only under the pool lock
    int32_t n;                       // lazy tasks in flight
    int32_t pad_[3];
    int32_t held[LAZY_MAX], total[LAZY_MAX], rpp[LAZY_MAX], slot[LAZY_MAX];   // pages held / needed, rows per page, CTA slot
    int32_t done[LAZY_MAX];          // scratch of the safety check
    float b_held[LAZY_BUCKETS], b_rate[LAZY_BUCKETS];                         // scratch of the projection
};
struct TbPool {
    uint8_t* base;          // n_pages * page_bytes
    int64_t page_bytes;
    int32_t n_pages;
    int32_t* free_stack;    // page ids, free_stack[0 .. *n_free)
    int32_t* n_free;
    int32_t* lock;          // 0 = free
    int32_t* progress;      // bumped (under the lock) by every grant and every release: the stall watchdog's heartbeat
    int32_t* gate;          // waiters for memory poll ONE at a time (the checks run under the pool lock)
    int32_t* reserve;       // pages a segmented task is waiting for (its whole traceback, taken at once): no NEW ordinary task starts
                            // unless that many stay free (running tasks still grow, so nothing can deadlock on it)
    LazyState* lazy;        // null = every task takes its pages up front
    int32_t* slot_idx;      // per CTA slot: index into lazy->* of the task it runs, or -1
    int32_t lazy_min_pages; // tasks with at least this many pages grow lazily (<= 0: none)
    int32_t bucket_rows;    // antidiagonals per projection bucket
    float lazy_fill;        // projected peak of the lazy tasks must stay below lazy_fill * n_pages
    int32_t stall_ms;       // watchdog: no grant and no release on the whole device for this long = broken, trap
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

// Under the pool lock: can every lazy task still finish, one after the other, if `free_after` pages
// stay free?  Entry `self` (if >= 0) is evaluated as holding `self_held` pages.
__device__ inline bool lazy_safe(const TbPool& P, int free_after, int self, int self_held)
{
    volatile LazyState* L = P.lazy;
    const int n = L->n;
    for (int i = 0; i < n; ++i) L->done[i] = 0;
    int av = free_after, left = n;
    bool progress = true;
    while (left > 0 && progress) {
        progress = false;
        for (int i = 0; i < n; ++i) {
            if (L->done[i]) continue;
            const int h = i == self ? self_held : L->held[i];
            if (L->total[i] - h <= av) { av += h; L->done[i] = 1; --left; progress = true; }
        }
    }
    return left == 0;
}
// Under the pool lock: projected peak (pages) of the lazy tasks if a new one (total pages, rows per page) starts now.
__device__ inline float lazy_projected_peak(const TbPool& P, int new_total, int new_rpp)
{
    volatile LazyState* L = P.lazy;
    for (int b = 0; b < LAZY_BUCKETS; ++b) { L->b_held[b] = 0.f; L->b_rate[b] = 0.f; }
    const int n = L->n;
    for (int i = 0; i <= n; ++i) {
        const int h = i < n ? L->held[i] : 0, t = i < n ? L->total[i] : new_total, rp = i < n ? L->rpp[i] : new_rpp;
        int b = (int)(((long long)(t - h) * rp) / P.bucket_rows);       // bucket in which the task ends
        if (b >= LAZY_BUCKETS) b = LAZY_BUCKETS - 1;
        L->b_held[b] += (float)h; L->b_rate[b] += 1.0f / (float)rp;
    }
    float alive_h = 0.f, alive_r = 0.f, peak = 0.f;
    for (int b = LAZY_BUCKETS - 1; b >= 0; --b) {                        // tasks of bucket b count as alive until its end
        alive_h += L->b_held[b]; alive_r += L->b_rate[b];
        const float u = alive_h + alive_r * (float)(b + 1) * (float)P.bucket_rows;
        peak = u > peak ? u : peak;
    }
    return peak;
}

// All of these are called by ONE thread of a CTA.
// Up-front allocation of n pages; false when they are not free or would endanger the lazy tasks.
// `for_seg`: the caller is the one segmented task whose turn it is (it owns the reserve and ignores it).
__device__ inline bool pool_try_alloc(const TbPool& P, int n, int32_t* table, bool for_seg = false)
{
    if (n <= 0) return true;
    bool ok = false;
    pool_lock(P.lock);
    int nf = *(volatile int32_t*)P.n_free;
    const int keep = for_seg ? 0 : *(volatile int32_t*)P.reserve;
    if (nf - n >= keep && (!P.lazy || ((volatile LazyState*)P.lazy)->n == 0 || lazy_safe(P, nf - n, -1, 0))) {
        for (int i = 0; i < n; ++i) table[i] = ((volatile int32_t*)P.free_stack)[nf - 1 - i];
        *(volatile int32_t*)P.n_free = nf - n;
        ++*(volatile int32_t*)P.progress;
        ok = true;
    }
    pool_unlock(P.lock);
    return ok;
}
// Admission of a lazy task of `total` pages run by CTA slot `slot`: takes its first min(2, total) pages.
__device__ inline bool pool_lazy_admit(const TbPool& P, int slot, int total, int rpp, int32_t* table, int& held)
{
    bool ok = false;
    const int first = total < 2 ? total : 2;
    pool_lock(P.lock);
    volatile LazyState* L = P.lazy;
    const int nf = *(volatile int32_t*)P.n_free, n = L->n;
    if (nf - first >= *(volatile int32_t*)P.reserve && n < LAZY_MAX) {
        bool fits = n == 0 || lazy_projected_peak(P, total, rpp) <= P.lazy_fill * (float)P.n_pages;
        if (fits) {
            L->held[n] = first; L->total[n] = total; L->rpp[n] = rpp; L->slot[n] = slot; L->n = n + 1;
            if (lazy_safe(P, nf - first, -1, 0)) {
                for (int i = 0; i < first; ++i) table[i] = ((volatile int32_t*)P.free_stack)[nf - 1 - i];
                *(volatile int32_t*)P.n_free = nf - first;
                ++*(volatile int32_t*)P.progress;
                ((volatile int32_t*)P.slot_idx)[slot] = n;
                held = first; ok = true;
            } else L->n = n;
        }
    }
    pool_unlock(P.lock);
    return ok;
}
// One more page for the lazy task of CTA slot `slot` (table[held] receives it).
static __device__ __noinline__ bool pool_lazy_grab(const TbPool& P, int slot, int32_t* table, int& held)
{
    bool ok = false;
    pool_lock(P.lock);
    volatile LazyState* L = P.lazy;
    const int nf = *(volatile int32_t*)P.n_free, me = ((volatile int32_t*)P.slot_idx)[slot];
    if (nf >= 1 && lazy_safe(P, nf - 1, me, held + 1)) {
        table[held] = ((volatile int32_t*)P.free_stack)[nf - 1];
        *(volatile int32_t*)P.n_free = nf - 1;
        ++*(volatile int32_t*)P.progress;
        L->held[me] = ++held;
        ok = true;
    }
    pool_unlock(P.lock);
    return ok;
}
// Gives back table[0 .. n); `slot` >= 0 also retires the lazy task of that CTA slot.
__device__ inline void pool_free(const TbPool& P, int n, const int32_t* table, int slot = -1)
{
    if (n <= 0 && slot < 0) return;
    pool_lock(P.lock);
    int nf = *(volatile int32_t*)P.n_free;
    for (int i = 0; i < n; ++i) ((volatile int32_t*)P.free_stack)[nf + i] = table[i];
    *(volatile int32_t*)P.n_free = nf + n;
    ++*(volatile int32_t*)P.progress;
    if (slot >= 0) {
        volatile LazyState* L = P.lazy;
        const int me = ((volatile int32_t*)P.slot_idx)[slot], last = L->n - 1;
        if (me != last) {        // keep the list dense
            L->held[me] = L->held[last]; L->total[me] = L->total[last]; L->rpp[me] = L->rpp[last];
            const int s2 = L->slot[last]; L->slot[me] = s2; ((volatile int32_t*)P.slot_idx)[s2] = me;
        }
        L->n = last; ((volatile int32_t*)P.slot_idx)[slot] = -1;
    }
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

// ---------------------------------------------------------------------------
// Segmented tasks.  The difference recurrence forgets its start after about 2w antidiagonals (DESIGN.md section 3.7,
// scripts/convergence_probe.py), so a long task is cut into segments that run on different CTAs at the same time:
// segment s > 0 starts COLD `warm` antidiagonals before its first own row, writes traceback rows and one record per
// antidiagonal from its first own row on, and dumps its state at its first and last own rows.  The task's traceback
// pages come from the dynamic pool, all at once, when its first segment starts (tasks take their turn in queue order
// and the one waiting holds a reserve against new ordinary tasks), and go back when its CIGAR is written.  The CTA that
// finishes the task's last segment replays the ksw_extz_t bookkeeping over the records, checks the boundaries below the
// segment the alignment ends in bit for bit (the state a segment reached cold == the state its predecessor reached from
// the truth), and walks the CIGAR.  A boundary that does not hold is REPAIRED: the same CTA runs the upper segment again,
// this time from its predecessor's end state (exact by construction), and checks on from there.
struct DevSeg {
    int32_t task;               // index into tasks[]
    int32_t index, count;       // this segment, segments of the task
    int32_t r0;                 // first antidiagonal computed (cold start; 0 for segment 0)
    int32_t r_begin, r_end;     // antidiagonals this segment owns: [r_begin, r_end)
};
struct SegTask {
    int64_t rec_off;            // into seg_rec: one int4 per antidiagonal {max H, argmax column, H[en0] | NEG_INF, H[st0] | NEG_INF}
    int64_t snap_off;           // into seg_snap (words): boundary b = slots 2b (last row of segment b) and 2b+1 (warm end of b+1)
    int32_t table_off;          // into seg_tables: the task's page table (all rows; filled when the task is admitted)
    int32_t n_segs, first_seg;
    int32_t ticket;             // admission order of the segmented tasks (= their order in the segment queue)
};
constexpr int SEG_SNAP_HDR = 8;                 // words: anchor H (lane st0), valid flag, "repaired" flag (slot 2b+1: segment b+1 was run again
                                                // from segment b's end state, so boundary b holds by construction and both share one score frame)
constexpr int SEG_SNAP_PER_THREAD = 67;         // 64 state words, Vt, Hb, reserved
constexpr int SEG_SNAP_WORDS = SEG_SNAP_HDR + 256 * SEG_SNAP_PER_THREAD;

// everything a fill kernel needs besides its own parameters
struct RunCtx {
    const uint8_t* qarena;
    const uint8_t* tarena;
    const DevTask* tasks;
    fsv_result* results;
    TbPool pool;
    int32_t* page_tables;        // per CTA: max_pages_per_task entries
    int32_t slot_base;           // pool slot of this launch's CTA 0 (slot = slot_base + blockIdx.x)
    int32_t max_pages_per_task;
    uint32_t* cigar;             // compact CIGAR arena
    unsigned long long* cigar_cursor;
    int64_t cigar_cap;           // words
    int32_t* overflow;           // set when the arena is too small
    long long* timeline;         // per task (caller order): start ns, end ns (globaltimer), or null
    DevScoring sc;
    // segmented tasks (null / unused when the batch has none)
    const DevSeg* segs;
    const SegTask* seg_tasks;
    int4* seg_rec;
    uint32_t* seg_snap;
    const int32_t* seg_tables;
    int32_t* seg_done;           // per segmented task: segments finished (+ 1 << 20 per repaired segment, counted by the host)
    int32_t* seg_admitted;       // per segmented task: its traceback pages are in its table (set by the CTA that runs its segment 0)
    int32_t* seg_cancel;         // per segmented task: set by segment 0 when the alignment z-drops inside it; the other segments poll it and stop
    int32_t* seg_ticket;         // ticket of the segmented task whose turn it is to take its pages
    int32_t* seg_foot;           // per segment: antidiagonal at which the band ran out inside it, or -1
};

__device__ __forceinline__ long long global_ns()
{
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Stall watchdog of the lazy pool: a waiter traps (the launch fails, nothing hangs) if NO task on the device
// got or returned a page for `stall_ms` (a minute by default) — with lazy growth every running task does so many times a second.
struct StallWatch {
    long long t0 = 0; int32_t seen = 0;
    __device__ __forceinline__ void poll(const TbPool& P)
    {
        if (!P.lazy) return;
        const int32_t p = *(volatile int32_t*)P.progress;
        const long long now = global_ns();
        if (!t0 || p != seen) { t0 = now; seen = p; }
        else if (now - t0 > (long long)P.stall_ms * 1000000ll) {
            volatile LazyState* L = P.lazy;
            printf("fsv pool stall: block %d free %d lazy %d:", (int)blockIdx.x, *(volatile int32_t*)P.n_free, L->n);
            for (int i = 0; i < L->n && i < 24; ++i) printf(" [%d/%d s%d]", L->held[i], L->total[i], L->slot[i]);
            printf("\n");
            __trap();
        }
    }
};

// A CTA may KEEP the pages of the task it has just finished (table[0 .. kept)) for its next one: a batch of 10^5 small tasks (the
// fills of the chained path, reads against a window) otherwise takes and returns one page per task under the one pool lock, which
// then is what the batch waits for (2 x 10^5 lock hand-overs of a microsecond each).  Only short, up-front tasks; only while the pool
// is comfortably free and no segmented task is waiting for its reserve (pool_may_keep).
constexpr int POOL_KEEP_MAX = 8;
__device__ __forceinline__ bool pool_may_keep(const TbPool& P, int n)
{
    return n > 0 && n <= POOL_KEEP_MAX && *(volatile int32_t*)P.reserve == 0 && *(volatile int32_t*)P.n_free >= (P.n_pages >> 2);
}

// Pages for task t (thread 0 only): up front, or lazily (`lazy_ok`, DPX kernels) when it is long.
// `held` receives the pages taken now; held < tb_pages marks a lazy task.  `kept`: pages of the CTA's previous task still in table[].
__device__ inline bool task_pages(const RunCtx& C, int t, int32_t* table, bool lazy_ok, int& held, int& kept)
{
    const DevTask& T = C.tasks[t];
    const bool lazy = lazy_ok && C.pool.lazy && C.pool.lazy_min_pages > 0 && T.tb_pages >= C.pool.lazy_min_pages;
    if (kept > 0) {
        if (!lazy && T.tb_pages >= kept && (T.tb_pages == kept || pool_try_alloc(C.pool, T.tb_pages - kept, table + kept))) {
            atomicAdd(C.pool.progress, 1);           // the watchdog's heartbeat
            held = T.tb_pages; kept = 0;
            return true;
        }
        pool_free(C.pool, kept, table);              // of no use for this task (or the rest is not to be had right now)
        kept = 0;
    }
    if (lazy) return pool_lazy_admit(C.pool, C.slot_base + (int)blockIdx.x, T.tb_pages, T.rows_per_page, table, held);
    if (!pool_try_alloc(C.pool, T.tb_pages, table)) return false;
    held = T.tb_pages;
    return true;
}

// One attempt of a WAITING CTA: only one waiter at a time runs the (locked, not free) admission checks, the
// others sleep, so that the tasks that are running never queue behind a crowd of pollers for the pool lock.
__device__ inline bool task_pages_gated(const RunCtx& C, int t, int32_t* table, bool lazy_ok, int& held, int& kept)
{
    if (atomicCAS(C.pool.gate, 0, 1) != 0) return false;
    const bool ok = task_pages(C, t, table, lazy_ok, held, kept);
    __threadfence();
    atomicExch(C.pool.gate, 0);
    return ok;
}

// Picks the next task for this CTA (thread 0 only) and gets its traceback pages.
// `pending` carries a claimed task that is still waiting for memory.  Returns the task index or -1.
__device__ inline int next_task(const RunCtx& C, const TaskQueue& Q, int32_t* table, int& pending, bool lazy_ok, int& held, int& kept)
{
    StallWatch watch;
    for (;;) {
        if (pending >= 0) {
            if (task_pages_gated(C, pending, table, lazy_ok, held, kept)) { int t = pending; pending = -1; return t; }
            int t = queue_take(Q, 1);                       // memory is short: do a small task meanwhile
            if (t >= 0) {
                if (task_pages(C, t, table, lazy_ok, held, kept)) return t;
                // not even the small one fits right now: wait for running tasks to finish
                for (;;) {
                    __nanosleep(20000);
                    if (task_pages_gated(C, t, table, lazy_ok, held, kept)) return t;
                    watch.poll(C.pool);
                }
            }
            __nanosleep(20000);                             // queue drained: wait for memory
            watch.poll(C.pool);
            continue;
        }
        int t = queue_take(Q, 0);
        if (t < 0) {
            if (kept > 0) { pool_free(C.pool, kept, table); kept = 0; }      // the queue is empty: nothing to keep pages for
            return -1;
        }
        if (task_pages(C, t, table, lazy_ok, held, kept)) return t;
        pending = t;
    }
}

// A lazily growing task keeps one page ahead of the antidiagonal it is writing (thread 0 only; spins while the
// grant would not be safe).  Out of line: this is rare code that must not cost the fill loop registers.
static __device__ __noinline__ int pool_lazy_grow(const TbPool& P, int slot, int32_t* table, int held, int want)
{
    StallWatch watch;
    while (held < want)
        if (!pool_lazy_grab(P, slot, table, held)) { __nanosleep(4000); watch.poll(P); }
    return held;
}

// byte address of traceback row r of a task (rows_per_page = page_bytes / pitch)
__device__ __forceinline__ uint8_t* tb_row(const TbPool& P, const int32_t* table, int rows_per_page, int pitch, int r)
{
    const int pg = r / rows_per_page;
    // (__ldcg: a segmented task's table is written by the CTA that admitted it, possibly on another SM, and L1 lines are shared by tables)
    return P.base + (int64_t)__ldcg(table + pg) * P.page_bytes + (int64_t)(r - pg * rows_per_page) * pitch;
}

}  // namespace fsv
