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
