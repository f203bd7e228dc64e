// fsv_backtrack.cuh — CIGAR reconstruction from the packed traceback rows.
//
// Semantics are those of ksw_backtrack / ksw_push_cigar
// (software/hifiasm-0.16.1/ksw2.h:103-151, is_rot = 1, min_intron_len = 0):
// a five-state walk {H, E, F, E~, F~} from the end cell towards (0,0), forced
// states outside the stored band, leading D / I, optional reversal.
//
// B200 shape: the walk runs in the fill kernel's own CTA right after the last
// antidiagonal (the rows it reads first are the ones written last, still in L2), on
// ONE warp.  It is a pointer chase, so instead of one dependent load per step the
// 32 lanes speculatively fetch the next 32 x BT_DEPTH cells along the direction the
// current state moves in (diagonal for H, column for E/E~, row for F/F~), every lane
// evaluates the state machine for "its" cells assuming the run continues, and the
// ballots find where the run really ends: a run of n equal steps costs
// n / (32 x BT_DEPTH) memory round trips.  off[r] / off_end[r] of the reference are recomputed from r
// (band_limits), so no per-row arrays are stored.  Two passes: count, then write into
// an exactly-sized slice of the compact CIGAR arena (one atomicAdd per task).
#pragma once
#include "fsv_common.cuh"

namespace fsv {

// next state at a cell reached in state `s` (ksw2.h:133-136)
__device__ __forceinline__ int bt_next_state(int s, int cell, int force)
{
    int ns = s;
    if (s == 0) ns = cell & 7;
    else if (!((cell >> (s + 2)) & 1)) ns = cell & 7;
    if (force >= 0) ns = force;
    return ns;
}

// All 32 lanes of one warp call this.  Returns the number of CIGAR words; with WRITE it also stores
// them (BAM encoding, len << 4 | op) at out[0 .. total).
constexpr int BT_DEPTH = 4;      // cells each lane fetches per round of the speculative walk

template <bool WRITE>
__device__ int bt_walk(const TbPool& pool, const int32_t* table, const DevTask& T, int i0, int j0, uint32_t* out, int total)
{
    const int lane = threadIdx.x & 31;
    const bool keep_rev = (T.flag & FSV_EZ_REV_CIGAR) != 0;
    int i = i0, j = j0, state = 0;
    int cur_op = -1, cur_len = 0, n_out = 0;
    auto flush = [&]() {
        if (cur_len > 0) {
            if (WRITE && lane == 0) {
                const int pos = keep_rev ? n_out : total - 1 - n_out;
                out[pos] = (uint32_t)cur_len << 4 | (uint32_t)cur_op;
            }
            ++n_out;
        }
    };
    auto emit = [&](int op, int n) {
        if (op == cur_op) cur_len += n;
        else { flush(); cur_op = op; cur_len = n; }
    };
    auto op_of = [](int s) { return s == 0 ? 0 : ((s == 1 || s == 3) ? 2 : 1); };

    while (i >= 0 && j >= 0) {
        // lane l looks at the cells l, l + 32, ... (BT_DEPTH of them) steps further along the direction of `state`: the loads of
        // one round are independent, so a run of n equal steps costs n / (32 * BT_DEPTH) memory round trips
        const int di = (state == 0 || state == 1 || state == 3) ? 1 : 0;
        const int dj = (state == 0 || state == 2 || state == 4) ? 1 : 0;
        int ns[BT_DEPTH]; bool valid[BT_DEPTH];
#pragma unroll
        for (int k = 0; k < BT_DEPTH; ++k) {
            const int step = lane + 32 * k;
            const int li = i - di * step, lj = j - dj * step;
            valid[k] = li >= 0 && lj >= 0;
            ns[k] = -1;
            if (valid[k]) {
                int r = li + lj, st0, en0;
                band_limits(r, T.qlen, T.tlen, T.w, st0, en0);
                const int st = round_st(st0), en = round_en(en0);
                int force = -1;
                if (li < st) force = 2;
                if (li > en) force = 1;
                int cell = 0;
                if (force < 0) {
                    // (.cg: the rows of a segmented task were written on other SMs, into pages this SM may have read for an earlier task)
                    cell = __ldcg(tb_row(pool, table, T.rows_per_page, T.pitch, r) + (li - st));
                    // the DPX kernel stores the winner's priority code (tb_mode - d) in bits 0-2
                    if (T.tb_mode) cell = (cell & ~7) | (T.tb_mode - (cell & 7));
                }
                ns[k] = bt_next_state(state, cell, force);
            }
        }
        int n = 0, ns_n = -1; bool valid_n = false, ended = false;
#pragma unroll
        for (int k = 0; k < BT_DEPTH; ++k) {
            const unsigned cont = __ballot_sync(0xffffffffu, valid[k] && ns[k] == state);
            int nk = __ffs(~cont) - 1;        // leading lanes of this group that stay in `state`
            if (nk < 0) nk = 32;
            const int src = nk < 32 ? nk : 0;
            const int ns_k = __shfl_sync(0xffffffffu, ns[k], src);
            const bool valid_k = __shfl_sync(0xffffffffu, (int)valid[k], src) != 0;
            if (!ended) {
                n += nk;
                if (nk < 32) { ended = true; ns_n = ns_k; valid_n = valid_k; }
            }
        }
        if (n > 0) { emit(op_of(state), n); i -= di * n; j -= dj * n; }
        if (ended) {
            if (!valid_n) break;             // walked off the matrix
            state = ns_n;
            if (state > 4) break;            // (not a ksw2 state: only a corrupted traceback row could hold it; never loop on it)
            emit(op_of(state), 1);
            if (state == 0) { --i; --j; }
            else if (state == 1 || state == 3) --i;
            else --j;
        }
    }
    if (i >= 0) emit(2, i + 1);              // leading deletion (ksw2.h:145)
    if (j >= 0) emit(1, j + 1);              // leading insertion (ksw2.h:146)
    flush();
    return n_out;
}

// End of a task, executed by warp 0 of the CTA after a CTA barrier: end-point choice
// (ksw2_extz2_sse.c:292-301), CIGAR into the arena, result record.
__device__ inline void finish_task(const RunCtx& C, const DevTask& T, const int32_t* table, const EzState& ez,
                                   int64_t cells, bool with_tb)
{
    const int lane = threadIdx.x & 31;
    int i0 = -1, j0 = -1, reach_end = 0;
    if (with_tb) {
        if (!ez.zdropped && !(T.flag & FSV_EZ_EXTZ_ONLY)) { i0 = T.tlen - 1; j0 = T.qlen - 1; }
        else if (!ez.zdropped && (T.flag & FSV_EZ_EXTZ_ONLY) && ez.mqe + T.end_bonus > ez.max) {
            reach_end = 1; i0 = ez.mqe_t; j0 = T.qlen - 1;
        } else if (ez.max_t >= 0 && ez.max_q >= 0) { i0 = ez.max_t; j0 = ez.max_q; }
    }
    int n_cigar = 0;
    long long off = 0;
    if (i0 >= 0 && j0 >= 0) {
        n_cigar = bt_walk<false>(C.pool, table, T, i0, j0, nullptr, 0);
        if (lane == 0) off = (long long)atomicAdd(C.cigar_cursor, (unsigned long long)n_cigar);
        off = __shfl_sync(0xffffffffu, off, 0);
        if (off + n_cigar <= C.cigar_cap) bt_walk<true>(C.pool, table, T, i0, j0, C.cigar + off, n_cigar);
        else if (lane == 0) atomicExch(C.overflow, 1);
    }
    if (lane == 0) {
        fsv_result R;
        R.max = ez.max; R.zdropped = ez.zdropped; R.max_q = ez.max_q; R.max_t = ez.max_t;
        R.mqe = ez.mqe; R.mqe_t = ez.mqe_t; R.mte = ez.mte; R.mte_q = ez.mte_q; R.score = ez.score;
        R.reach_end = reach_end; R.n_cigar = n_cigar; R.status = 0; R.cigar_off = off; R.cells = cells;
        C.results[T.orig] = R;
    }
}

__device__ inline void finish_reset_task(const RunCtx& C, const DevTask& T)
{   // ksw2's silent returns (ksw2_extz2_sse.c:57,82): a successful task with a reset result
    fsv_result R;
    R.max = 0; R.zdropped = 0; R.max_q = R.max_t = R.mqe_t = R.mte_q = -1;
    R.mqe = R.mte = R.score = FSV_NEG_INF; R.reach_end = 0; R.n_cigar = 0;
    R.status = T.pad_; R.cigar_off = 0; R.cells = 0;
    C.results[T.orig] = R;
}

}  // namespace fsv
