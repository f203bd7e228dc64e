// fsv_backtrack.cuh — CIGAR reconstruction from the packed traceback rows.
//
// Semantics are those of ksw_backtrack / ksw_push_cigar
// (software/hifiasm-0.16.1/ksw2.h:103-151, is_rot = 1, min_intron_len = 0):
// a five-state walk {H, E, F, E~, F~} from the end cell towards (0,0), forced
// states outside the stored band, leading D / I, optional reversal.
//
// B200 shape: one warp per task.  The walk is a pointer chase through HBM, so
// instead of one dependent load per step the 32 lanes speculatively fetch the
// next 32 cells along the direction the current state moves in (diagonal for H,
// column for E/E~, row for F/F~), every lane evaluates the state machine for
// "its" cell assuming the run continues, and one ballot finds where the run
// really ends.  A run of n equal steps costs one memory round trip.
// off[r] / off_end[r] of the reference are recomputed from r (band_limits), so
// no per-row arrays are stored.
#pragma once
#include "fsv_common.cuh"

namespace fsv {

constexpr int BT_THREADS = 128;   // 4 tasks per CTA

struct BtParams {
    const DevTask* tasks;
    const int32_t* order;
    int32_t n_order;
    fsv_result* results;
    DevAux* aux;
    const uint8_t* tb;
    uint32_t* cigar;          // compact CIGAR arena (write pass)
    int64_t cigar_cap;        // words
    int32_t* overflow;        // set when a CIGAR does not fit
};

// next state at a cell reached in state `s` (ksw2.h:133-136)
__device__ __forceinline__ int bt_next_state(int s, int cell, int force)
{
    int ns = s;
    if (s == 0) ns = cell & 7;
    else if (!((cell >> (s + 2)) & 1)) ns = cell & 7;
    if (force >= 0) ns = force;
    return ns;
}

template <bool WRITE>
__global__ void __launch_bounds__(BT_THREADS) fsv_backtrack_kernel(const BtParams P)
{
    const int wslot = (blockIdx.x * BT_THREADS + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wslot >= P.n_order) return;
    const int ti = P.order[wslot];
    const DevTask T = P.tasks[ti];
    const DevAux A = P.aux[ti];
    if (T.kind == 0 || A.i0 < 0 || A.j0 < 0 || T.tb_off < 0) {
        if (!WRITE && lane == 0) { P.aux[ti].n_cigar = 0; P.results[T.orig].n_cigar = 0; }
        return;
    }
    const uint8_t* tb = P.tb + T.tb_off;
    const bool keep_rev = (T.flag & FSV_EZ_REV_CIGAR) != 0;
    const int total = A.n_cigar;                        // valid in the write pass
    const int64_t out_base = WRITE ? P.results[T.orig].cigar_off : 0;
    const bool fits = WRITE ? (out_base + total <= P.cigar_cap) : false;

    int i = A.i0, j = A.j0, state = 0;
    int cur_op = -1, cur_len = 0, n_out = 0;
    auto flush = [&]() {
        if (cur_len > 0) {
            if (WRITE && fits && lane == 0) {
                int pos = keep_rev ? n_out : total - 1 - n_out;
                P.cigar[out_base + pos] = (uint32_t)cur_len << 4 | (uint32_t)cur_op;
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
        // lane l looks at the cell l steps further along the direction of `state`
        int di = (state == 0 || state == 1 || state == 3) ? 1 : 0;
        int dj = (state == 0 || state == 2 || state == 4) ? 1 : 0;
        int li = i - di * lane, lj = j - dj * lane;
        bool valid = li >= 0 && lj >= 0;
        int ns = -1;
        if (valid) {
            int r = li + lj, st0, en0;
            band_limits(r, T.qlen, T.tlen, T.w, st0, en0);
            int st = round_st(st0), en = round_en(en0);
            int force = -1;
            if (li < st) force = 2;
            if (li > en) force = 1;
            int cell = force < 0 ? tb[(int64_t)r * T.pitch + (li - st)] : 0;
            // the DPX kernel stores the winner's priority code (tb_mode - d) in bits 0-2
            if (T.tb_mode && force < 0) cell = (cell & ~7) | (T.tb_mode - (cell & 7));
            ns = bt_next_state(state, cell, force);
        }
        unsigned cont = __ballot_sync(0xffffffffu, valid && ns == state);
        int n = __ffs(~cont) - 1;            // leading lanes that stay in `state`
        if (n < 0) n = 32;
        if (n > 0) { emit(op_of(state), n); i -= di * n; j -= dj * n; }
        if (n < 32) {
            int ns_n = __shfl_sync(0xffffffffu, ns, n);
            bool valid_n = __shfl_sync(0xffffffffu, (int)valid, n) != 0;
            if (!valid_n) break;             // walked off the matrix
            state = ns_n;
            emit(op_of(state), 1);
            if (state == 0) { --i; --j; }
            else if (state == 1 || state == 3) --i;
            else --j;
        }
    }
    if (i >= 0) emit(2, i + 1);              // leading deletion (ksw2.h:145)
    if (j >= 0) emit(1, j + 1);              // leading insertion (ksw2.h:146)
    flush();
    if (!WRITE) {
        if (lane == 0) { P.aux[ti].n_cigar = n_out; P.results[T.orig].n_cigar = n_out; }
    } else if (!fits && lane == 0) atomicExch(P.overflow, 1);
}

// exclusive scan of n_cigar over the tasks of one chunk, continuing from *running
// (single CTA; a chunk holds at most a few thousand tasks)
__global__ void fsv_cigar_offsets_kernel(const DevTask* tasks, const int32_t* order, int32_t n_order,
                                         const DevAux* aux, fsv_result* results, int64_t* running)
{
    __shared__ int64_t sh_warp[32];
    __shared__ int64_t sh_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sh_base = *running;
    __syncthreads();
    for (int base = 0; base < n_order; base += (int)blockDim.x) {
        int idx = base + tid;
        int64_t v = 0; int ti = -1;
        if (idx < n_order) { ti = order[idx]; v = aux[ti].n_cigar; }
        int64_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int64_t n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) sh_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int64_t wv = lane < (int)(blockDim.x >> 5) ? sh_warp[lane] : 0, wi = wv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int64_t n = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += n;
            }
            sh_warp[lane] = wi - wv;     // exclusive prefix of the warp totals
        }
        __syncthreads();
        int64_t excl = sh_base + sh_warp[warp] + incl - v;
        if (ti >= 0) results[tasks[ti].orig].cigar_off = excl;
        __syncthreads();
        if (tid == (int)blockDim.x - 1) sh_base = excl + v;
        __syncthreads();
    }
    if (tid == 0) *running = sh_base;
}

}  // namespace fsv
