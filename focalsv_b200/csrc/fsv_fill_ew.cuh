// fsv_fill_ew.cuh — the fill kernel with a BAND-EDGE WARP.
//
// fsv_fill_dpx_kernel (fsv_fill_dpx.cuh) keeps every 16-lane vector of the band in one thread, the vectors at the two band edges
// included.  Those two need everything that is irregular about ksw2's band (software/hifiasm-0.16.1/ksw2_extz2_sse.c:101-147,
// 224-269): carries into the first vector, the first row, profile stores that start at st0 and overhang en0, H[en0] re-derived
// from its left neighbour, maxima over the exact band only.  In packed form that is lane masks and selects for the whole warp
// the edge vectors happen to sit in: that warp needs 1.5x the time of the others per antidiagonal (ncu: 36 % of all warp time
// is spent waiting for it at the CTA barrier), and a band of w = 500 (33 vectors) leaves half of a 2-warp CTA idle.
//
// Here a task gets NWM "main" warps plus ONE edge warp:
//   * main threads hold only vectors STRICTLY INSIDE the band (st_ < V < en_), packed as before (one thread = one vector, state in
//     registers, dpx_cells), and run only the branch-free interior path;
//   * the edge warp holds the lowest vector of the band in lanes 0..15 and the highest in lanes 16..31, ONE LANE = ONE CELL, so
//     that every irregularity is a per-lane predicate (t >= st0, t == en0, t == r ...) instead of a mask over eight words; its
//     lanes 16..31 also carry the score profile of the vector ABOVE the band (the reference's profile stores overhang en0);
//   * a vector moves between the two forms through shared memory when it enters the band's interior (32 antidiagonals after it
//     became the highest vector) and when it becomes the lowest one: 70 words per 32 antidiagonals;
//   * the recurrence itself is the same instruction sequence in both forms (ew_cell = one word of dpx_cells).
// The kernel takes the mainstream tasks only: left-aligned traceback (no KSW_EZ_RIGHT / SCORE_ONLY / APPROX_MAX), no wildcard
// bases, not segmented, band >= 32, both sequences >= 64 bases; everything else stays with fsv_fill_dpx_kernel, which is also
// what this kernel is tested against (same oracle digests).
#pragma once
#include <cstdio>
#include <cstdlib>

#include "fsv_fill_dpx.cuh"

namespace fsv {

// resident CTAs per SM the register budget is tuned for ((NWM + 1) warps per CTA)
#ifndef FSV_EW_OCC6
#define FSV_EW_OCC6 2
#endif
#ifndef FSV_EW_OCC1
#define FSV_EW_OCC1 6
#endif
#ifndef FSV_EW_OCC2
#define FSV_EW_OCC2 4
#endif
template <int NWM> struct EwOcc { static constexpr int value = NWM == 1 ? FSV_EW_OCC1 : NWM == 2 ? FSV_EW_OCC2 : NWM == 4 ? 2 : NWM == 6 ? FSV_EW_OCC6 : 1; };

// One DP cell (a lane of the reference's vector, :26-47, :171-196) in the LOW half of every word: the instruction sequence of
// dpx_cells for one word.  Returns the traceback byte.
template <bool DUAL>
__device__ __forceinline__ uint32_t ew_cell(uint32_t& u, uint32_t& v, uint32_t& x, uint32_t& y, uint32_t& x2, uint32_t& y2, uint32_t s,
                                            uint32_t xt1, uint32_t vt1, uint32_t x2t1, const DpxK& K, uint32_t ALL1)
{
    using C = DpxConst<DUAL, false>;
    const uint32_t fE = both(C::cE), fF = both(C::cF), fE2 = both(C::cE2), fF2 = both(C::cF2);
    const uint32_t oE = both(0x08u | C::cE), oF = both(0x10u | C::cF), oE2 = both(0x20u | C::cE2), oF2 = both(0x40u | C::cF2);
    const uint32_t ut = u;
    const uint32_t a = __vadd2(xt1, vt1), b = __vadd2(y, ut);
    uint32_t a2 = 0, b2 = 0, zk, nz;
    if (DUAL) {
        a2 = __vadd2(x2t1, vt1);
        b2 = __vadd2(y2, ut);
        zk = __vimax3_s16x2(__vimax3_s16x2(s, a, b), a2, b2);
        nz = __vmaxs2(~zk | 0x00ff00ffu, K.nClamp);
    } else {
        const uint32_t t1 = __vmaxs2(s, a);
        zk = __vmaxs2(t1, b);
        nz = __vmaxu2(__vminu2(~t1 | 0x00ff00ffu, ~b | 0x00ff00ffu), K.nClamp);
    }
    u = not_fma(__vadd2(nz, vt1), ALL1);
    v = not_fma(__vadd2(nz, ut), ALL1);
    const uint32_t n1 = __vadd2(K.kQ1, nz);
    const uint32_t xa = __viaddmax_s16x2(a, n1, fE), ya = __viaddmax_s16x2(b, n1, fF);
    uint32_t fl;
    if (DUAL) {
        const uint32_t n2 = __vadd2(K.kQ21, nz);
        const uint32_t xa2 = __viaddmax_s16x2(a2, n2, fE2), ya2 = __viaddmax_s16x2(b2, n2, fF2);
        x = __vadd2(xa, K.nQE); y = __vadd2(ya, K.nQE);
        x2 = __vadd2(xa2, K.nQE2); y2 = __vadd2(ya2, K.nQE2);
        fl = __vmins2(xa, oE) + __vmins2(ya, oF) + __vmins2(xa2, oE2) + __vmins2(ya2, oF2);
    } else {
        x = xa; y = ya;
        fl = __vmins2(xa, oE) + __vmins2(ya, oF);
    }
    return ((fl & 0x00780078u) | (zk & 0x00070007u)) & 0xffu;
}

__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ uint32_t lds16(uint32_t a) { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory"); return v; }

template <bool DUAL, int NWM, bool EXCL = false>
__global__ void __launch_bounds__((NWM + 1) * 32, EwOcc<NWM>::value) fsv_fill_ew_kernel(const __grid_constant__ DpxParams P)
{
    constexpr int NT = NWM * 32;            // main threads; the edge warp is warp NWM
    constexpr int NSLOT = (NWM + 1 + 3) / 4 * 4;
    using KC = DpxConst<DUAL, false>;
    // shared block:
    //   EDGE  [2 parities][NWM] x 32 B : {x, v, x2, qw, H} of lane 15 of every main warp's last vector (as in fsv_fill_dpx_kernel)
    //   EWLO  [2] x 32 B               : the same of lane 15 of the band's lowest vector (edge warp -> main thread of vector st_+1)
    //   MHI   [2] x 32 B               : the same of lane 15 of vector en_-1 (its main thread -> lane 16 of the edge warp)
    //   MX    [3][NSLOT]               : per-warp maximum H of an antidiagonal (ring over 3); warp 0's slot = INT32_MAX means "stop"
    //   KEY / HEN0 / HST0 [3]          : tie key, H[en0], H[st0] rings; TASK, HELD
    //   LOST / HIST 72 words each      : a vector on its way main -> edge warp (becomes the lowest) / edge warp -> main (enters the interior):
    //                                    U V X Y X2 Y2 S Hr as 8 packed words each, then Hb, tw, qw
    constexpr uint32_t OFF_EDGE = 0, OFF_EWLO = OFF_EDGE + 2 * NWM * 32, OFF_MHI = OFF_EWLO + 64, OFF_MX = OFF_MHI + 64,
                       OFF_KEY = OFF_MX + 3 * NSLOT * 4, OFF_HEN0 = OFF_KEY + 16, OFF_HST0 = OFF_HEN0 + 16, OFF_TASK = OFF_HST0 + 16,
                       OFF_HELD = OFF_TASK + 4, OFF_LOST = OFF_HELD + 12, OFF_HIST = OFF_LOST + 72 * 4, SH_BYTES = OFF_HIST + 72 * 4;
    static_assert(OFF_LOST % 16 == 0 && OFF_MX % 16 == 0, "vector accesses");
    __shared__ __align__(16) uint32_t sh_raw[(SH_BYTES + 15) / 16 * 4];
    uint32_t sb = (uint32_t)__cvta_generic_to_shared(sh_raw);
    const RunCtx& C = P.C;
    const DevScoring& sc = C.sc;
    int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    asm volatile("" : "+r"(tid), "+r"(lane), "+r"(warp), "+r"(sb));
    const unsigned FULL = 0xffffffffu;
    const DpxK& K = P.K;
    uint32_t ALL1 = 0xffffffffu;
    asm volatile("" : "+r"(ALL1));
    const uint32_t extSel = DUAL ? 0xB391u : 0x4341u;
    const bool is_ew = warp == NWM;
    int32_t* table = C.page_tables + (int64_t)blockIdx.x * C.max_pages_per_task;
    int pending = -1, kept = 0;
    if (tid < 3 * NSLOT) sts32(sb + OFF_MX + 4u * (uint32_t)tid, (uint32_t)INT32_MIN);       // slots of warps that do not exist stay at the minimum

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            int held = 0;
            sts32(sb + OFF_TASK, (uint32_t)next_task(C, P.Q, table, pending, true, held, kept));
            sts32(sb + OFF_HELD, (uint32_t)held);
            // a "stop" of the previous task must not be seen by this one; keys start empty
            for (int j = 0; j < 3; ++j) { sts32(sb + OFF_MX + 4u * (uint32_t)(j * NSLOT), (uint32_t)INT32_MIN); sts32(sb + OFF_KEY + 4u * (uint32_t)j, 0xffffffffu); }
        }
        __syncthreads();
        const int ti = (int)lds32(sb + OFF_TASK);
        if (ti < 0) return;
        const DevTask T = C.tasks[ti];
        if (tid == 0 && C.timeline) C.timeline[2 * T.orig] = global_ns();
        const int qlen = T.qlen, tlen = T.tlen, w = T.w;
        const uint8_t* query = C.qarena + T.q_off;
        const uint8_t* target = C.tarena + T.t_off;
        int tb_rip = 0, tb_pg = 0;
        uint8_t* tb_page = C.pool.base + (int64_t)table[0] * C.pool.page_bytes;
        const int n_diag = qlen + tlen - 1;

        // ---- main threads: one interior vector each (or none yet / none any more)
        uint32_t U[8], V[8], X[8], Y[8], X2[8], Y2[8], S[8], Hr[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { U[k] = V[k] = X[k] = Y[k] = X2[k] = Y2[k] = S[k] = Hr[k] = 0; }
        int32_t Hb = 0;
        int Vt = tid == 0 ? NT : tid;          // the vector this thread holds, or the next one it will get (vector V lives in thread V mod NT; vector 0 never leaves the edge warp)
        bool has_vec = false;
        uint32_t tw = 0, qw = 0;
        // ---- edge warp: one cell per lane (low halves), group 0 = lanes 0..15 = vector st_, group 1 = lanes 16..31 = vector en_ (if en_ > st_);
        // *_ov: score profile / bases of vector en_+1 in the lanes of group 1
        const int c16 = lane & 15, grp = lane >> 4;
        // (the edge warp's state lives in the registers that hold a packed vector in a main thread: the two sets are never needed together)
        uint32_t &eu = U[0], &ev = V[0], &ex = X[0], &ey = Y[0], &ex2 = X2[0], &ey2 = Y2[0], &es = S[0];
        int32_t eH = 0;
        uint32_t &etb = U[1], &eqb = U[2], &s_ov = U[3], &tb_ov = U[4], &qb_ov = U[5];
        uint32_t &cx = V[1], &cv = V[2], &cx2 = V[3];            // lane 15 of the vector that just left the band at the bottom (the "inl" carry, :118-122) ...
        int32_t ch = 0;                                          // ... and its H
        uint32_t &nbx = X[1], &nbv = X[2], &nbx2 = X[3], &nbq = X[4];   // (lane 16) lane 15 of a vector that has just left group 1 for the interior
        int32_t nbh = 0;
        bool nb_saved = false, load_lo = false;
        int32_t hk = 0;                                        // H[en0-1] as the reference's H[] holds it (hprev_keep of fsv_fill_dpx_kernel)
        uint32_t &q_lo = Y[1];                                 // raw query byte lane 0 needs at the next antidiagonal
        int tp = -1; int32_t Hp = INT32_MIN;                   // this lane at r-1: column (or -1 = not inside the band) and H
        if (is_ew) {
            eu = K.gU & 0xffffu; ev = K.gU & 0xffffu; ex = K.gX & 0xffffu; ey = K.gY & 0xffffu; ex2 = K.gX2 & 0xffffu; ey2 = K.gY2 & 0xffffu; es = K.sInit & 0xffffu;
            s_ov = K.sInit & 0xffffu;
        }
        if (is_ew) {
            // antidiagonal 0: the band is cell (0, 0) of vector 0; vector 1 is the one above
            const int t0 = c16, t1 = 16 + c16;
            if (grp == 0) { etb = t0 < tlen ? (uint32_t)target[t0] & 3u : 0u; }
            else { tb_ov = t1 < tlen ? (uint32_t)target[t1] & 3u : 0u; }
            // query bases as they stand BEFORE antidiagonal 0 (lane c faces query[-1 - t]: nothing), so that the shift of antidiagonal 0 brings query[0] to lane 0
            q_lo = (uint32_t)query[0];
        }
        EzState ez; ez.reset();
        int64_t cells = 0;
        int stop_r = n_diag;
        bool dropped = false;
        int32_t maxrun = 0, M1 = 0, M2 = 0;
        bool nt1 = false, nt2 = false;
        int32_t habs_p = INT32_MIN; bool act_p = false; int Vp = 0;      // this thread at r-1: best H, did it compute, which vector
        int st0p = 0, en0p = 0, st_p = 0, en_p = 0;
        bool grad_p = false; int grad_v = -1;                  // a vector left the edge warp for the interior at the end of r-1
        int s3 = 0;
        int st0 = 0, en0 = 0;                                  // band of the antidiagonal about to be computed (computed one iteration ahead)
        band_limits(0, qlen, tlen, w, st0, en0);

        for (int r = 0;; ++r) {
            const int par = r & 1, ppar = par ^ 1;
            const int s3m1 = s3 == 0 ? 2 : s3 - 1, s3m2 = s3 == 2 ? 0 : s3 + 1;
            // ---- neighbour's lane 15 as it stood after the previous antidiagonal (main threads, packed: the value sits in the HIGH half)
            uint32_t nbX = 0, nbV = 0, nbX2 = 0, nbQ = 0;
            if (!is_ew) {
                nbX = __shfl_up_sync(FULL, X[7], 1); nbV = __shfl_up_sync(FULL, V[7], 1);
                nbX2 = DUAL ? __shfl_up_sync(FULL, X2[7], 1) : 0;
                nbQ = __shfl_up_sync(FULL, qw, 1);
                if (NWM == 1) {
                    const uint32_t a = __shfl_sync(FULL, X[7], 31), b = __shfl_sync(FULL, V[7], 31);
                    const uint32_t c2 = DUAL ? __shfl_sync(FULL, X2[7], 31) : 0, q2 = __shfl_sync(FULL, qw, 31);
                    if (lane == 0) { nbX = a; nbV = b; nbX2 = c2; nbQ = q2; }
                } else if (lane == 0 && r > 0) {
                    const uint4 e = lds128(sb + OFF_EDGE + (uint32_t)(ppar * NWM + (warp + NWM - 1) % NWM) * 32u);
                    nbX = e.x; nbV = e.y; nbX2 = e.z; nbQ = e.w;
                }
                // the vector right above the band's lowest one takes its neighbour from the edge warp (if that vector was the lowest at r-1 too)
                if (r > 0 && Vt - 1 == st_p) {
                    const uint4 e = lds128(sb + OFF_EWLO + (uint32_t)ppar * 32u);
                    nbX = e.x; nbV = e.y; nbX2 = e.z; nbQ = e.w;
                }
                // a vector that entered the interior at the end of r-1 arrives from the edge warp
                if (grad_p && !has_vec && Vt == grad_v) {
                    const uint32_t a = sb + OFF_HIST;
#pragma unroll
                    for (int q4 = 0; q4 < 2; ++q4) {
                        uint4 e;
                        e = lds128(a + 0 * 32 + q4 * 16); U[4 * q4] = e.x; U[4 * q4 + 1] = e.y; U[4 * q4 + 2] = e.z; U[4 * q4 + 3] = e.w;
                        e = lds128(a + 1 * 32 + q4 * 16); V[4 * q4] = e.x; V[4 * q4 + 1] = e.y; V[4 * q4 + 2] = e.z; V[4 * q4 + 3] = e.w;
                        e = lds128(a + 2 * 32 + q4 * 16); X[4 * q4] = e.x; X[4 * q4 + 1] = e.y; X[4 * q4 + 2] = e.z; X[4 * q4 + 3] = e.w;
                        e = lds128(a + 3 * 32 + q4 * 16); Y[4 * q4] = e.x; Y[4 * q4 + 1] = e.y; Y[4 * q4 + 2] = e.z; Y[4 * q4 + 3] = e.w;
                        e = lds128(a + 4 * 32 + q4 * 16); X2[4 * q4] = e.x; X2[4 * q4 + 1] = e.y; X2[4 * q4 + 2] = e.z; X2[4 * q4 + 3] = e.w;
                        e = lds128(a + 5 * 32 + q4 * 16); Y2[4 * q4] = e.x; Y2[4 * q4 + 1] = e.y; Y2[4 * q4 + 2] = e.z; Y2[4 * q4 + 3] = e.w;
                        e = lds128(a + 6 * 32 + q4 * 16); S[4 * q4] = e.x; S[4 * q4 + 1] = e.y; S[4 * q4 + 2] = e.z; S[4 * q4 + 3] = e.w;
                        e = lds128(a + 7 * 32 + q4 * 16); Hr[4 * q4] = e.x; Hr[4 * q4 + 1] = e.y; Hr[4 * q4 + 2] = e.z; Hr[4 * q4 + 3] = e.w;
                    }
                    const uint4 e = lds128(a + 8 * 32);
                    Hb = (int32_t)e.x; tw = e.y; qw = e.z;
                    has_vec = true;
                    // its neighbour below is interior too and was computed by thread tid-1: the shuffles above hold
                }
            }
            // maximum of antidiagonal r-1 (every warp's slot, written behind the last barrier), fetched early
            int32_t m_prev = INT32_MIN;
            if (r >= 1) {
                const uint32_t a = sb + OFF_MX + 4u * (uint32_t)(s3m1 * NSLOT);
                const uint4 p = lds128(a);
                m_prev = max(max((int32_t)p.x, (int32_t)p.y), max((int32_t)p.z, (int32_t)p.w));
                if (NSLOT > 4) { const uint4 q = lds128(a + 16u); m_prev = max(m_prev, max(max((int32_t)q.x, (int32_t)q.y), max((int32_t)q.z, (int32_t)q.w))); }
                if (NSLOT > 8) { const uint4 q = lds128(a + 32u); m_prev = max(m_prev, max(max((int32_t)q.x, (int32_t)q.y), max((int32_t)q.z, (int32_t)q.w))); }
            }
            // ---- (A) antidiagonal d = r-2 is final: bookkeeping (ksw2_extz2_sse.c:262-269), by warp 0
            if (r >= 2) {
                if (m_prev == INT32_MAX) { dropped = true; break; }
                maxrun = max(maxrun, M2);
                const int d = r - 2;
                if (warp == 0 && !dropped) {
                    int st0d, en0d;
                    band_limits(d, qlen, tlen, w, st0d, en0d);
                    cells += en0d - st0d + 1;
                    int max_t = en0d;
                    if (nt2) {
                        const uint32_t bk = lds32(sb + OFF_KEY + 4u * (uint32_t)s3m2);
                        if (bk != 0) max_t = (int)((bk - 1u) & ((1u << 26) - 1u));
                    }
                    int32_t h_last = FSV_NEG_INF;
                    if (en0d == tlen - 1) {
                        h_last = (int32_t)lds32(sb + OFF_HEN0 + 4u * (uint32_t)s3m2);
                        if (h_last > ez.mte) { ez.mte = h_last; ez.mte_q = d - round_en(en0d); }
                    }
                    if (d - st0d == qlen - 1) {
                        const int32_t h = (int32_t)lds32(sb + OFF_HST0 + 4u * (uint32_t)s3m2);
                        if (h > ez.mqe) { ez.mqe = h; ez.mqe_t = st0d; }
                    }
                    if (ez.apply_zdrop(M2, d, max_t, T.zdrop, sc.e_drop)) { dropped = true; if (lane == 0) sts32(sb + OFF_MX + 4u * (uint32_t)(s3 * NSLOT), (uint32_t)INT32_MAX); }
                    else if (d == n_diag - 1 && en0d == tlen - 1) ez.score = h_last;
                }
                if (d == stop_r - 1) {
                    __syncthreads();
                    if ((int32_t)lds32(sb + OFF_MX + 4u * (uint32_t)(s3 * NSLOT)) == INT32_MAX) dropped = true;
                    break;
                }
            }
            // ---- (B) antidiagonal r-1: its maximum, and (only if observable, ksw2.h:164-174) who holds it
            if (r >= 1 && r - 1 < stop_r) {
                const int32_t m = m_prev;
                M1 = m;
                const int32_t mr = max(maxrun, M2);
                nt1 = m > mr || (T.zdrop >= 0 && mr - m > T.zdrop);
                if (nt1) {
                    uint32_t key = 0xffffffffu;
                    const int en1 = st0p + (en0p - st0p) / 4 * 4;
                    if (is_ew) {
                        if (tp >= 0 && Hp == m) key = tp == en0p ? 0u : tp < en1 ? 1u + ((uint32_t)((tp - st0p) & 3) << 26) + (uint32_t)tp : 1u + (4u << 26) + (uint32_t)tp;
                    } else if (act_p && habs_p == m) {
                        // (an interior vector: all 16 lanes were inside the band, none of them is en0)
                        const int base = Vp << 4;
                        const uint32_t mv = both((uint32_t)(m - Hb) & 0xffffu);
                        uint32_t ne = 0;
#pragma unroll
                        for (int k = 0; k < 8; ++k) ne += __vminu2(Hr[k] ^ mv, 0x00010001u) << k;
                        const uint32_t eq = ~((ne & 0xffu) | ((ne >> 8) & 0xff00u)) & 0xffffu;
                        const int nb = min(max(en1 - base, 0), 16);
                        const uint32_t body = eq & ((1u << nb) - 1u), tail = eq & ~((1u << nb) - 1u);
                        const int sft = (base - st0p) & 3;
#pragma unroll
                        for (int i = 3; i >= 0; --i) {
                            const uint32_t mm = body & (0x1111u << ((i - sft) & 3));
                            if (mm) key = 1u + ((uint32_t)i << 26) + (uint32_t)(base + __ffs(mm) - 1);
                        }
                        if (!body && tail) key = 1u + (4u << 26) + (uint32_t)(base + __ffs(tail) - 1);
                    }
                    if (key != 0xffffffffu) atom_min_shared(sb + OFF_KEY + 4u * (uint32_t)s3m1, key);
                }
            }
            if (tid == 0) sts32(sb + OFF_KEY + 4u * (uint32_t)s3, 0xffffffffu);

            // ---- (C) compute antidiagonal r
            bool valid = false;
            if (r < stop_r) {
                if (st0 > en0) stop_r = r; else valid = true;
            }
            act_p = false; tp = -1;
            int st0n = st0, en0n = en0;
            if (valid) {
                const int st = st0 & ~15, st_ = st0 >> 4, en_ = en0 >> 4, en = en0 | 15;
                band_limits(r + 1, qlen, tlen, w, st0n, en0n);
                const bool nvalid = r + 1 < n_diag && st0n <= en0n;
                const int st_n = nvalid ? st0n >> 4 : st_, en_n = nvalid ? en0n >> 4 : en_;
                const bool two = en_ > st_;
                int32_t habs = INT32_MIN;
                if (!is_ew) {
                    if (has_vec) {
                        const int base = Vt << 4;
                        qw = (qw << 2) | (nbQ >> 30);
                        dpx_profile<DUAL, false>(S, tw, qw, K);
                        const uint32_t XT0 = prmt(nbX, X[7], 0x5432u), VT0 = prmt(nbV, V[7], 0x5432u);
                        const uint32_t X2T0 = DUAL ? prmt(nbX2, X2[7], 0x5432u) : 0;
                        uint4 o;
                        dpx_cells<DUAL, true, false>(U, V, X, Y, X2, Y2, S, XT0, VT0, X2T0, K, ALL1, o);
                        *reinterpret_cast<uint4*>(tb_page + (int64_t)tb_rip * T.pitch + (base - st)) = o;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            uint32_t dv = prmt(V[k], 0u, extSel);
                            if (!DUAL) dv = __vadd2(dv, K.nBias);
                            Hr[k] = __vadd2(Hr[k], dv);
                        }
                        const uint32_t pm = __vimax3_s16x2(__vimax3_s16x2(Hr[0], Hr[1], Hr[2]), __vimax3_s16x2(Hr[3], Hr[4], Hr[5]), __vmaxs2(Hr[6], Hr[7]));
                        const int mrel = max(sext16(pm), sext16(pm >> 16));
                        habs = Hb + mrel;
                        if (__builtin_expect((r & 31) == 31, 0)) {
                            Hb += mrel;
                            const uint32_t dd = both((uint32_t)mrel & 0xffffu);
#pragma unroll
                            for (int k = 0; k < 8; ++k) Hr[k] = __vsub2(Hr[k], dd);
                        }
                        act_p = true; Vp = Vt;
                        // lane 15 of the vector right below the band's highest one (at r+1), for lane 16 of the edge warp
                        if (Vt == en_n - 1 && Vt > st_n) {
                            const uint32_t ea = sb + OFF_MHI + (uint32_t)par * 32u;
                            sts128(ea, make_uint4(X[7], V[7], DUAL ? X2[7] : 0u, qw));
                            sts32(ea + 16u, (uint32_t)(Hb + sext16(Hr[7] >> 16)));
                        }
                        // this vector is the band's lowest one from r+1 on: it moves to the edge warp
                        if (__builtin_expect(Vt == st_n && nvalid, 0)) {
                            const uint32_t a = sb + OFF_LOST;
                            sts128(a + 0 * 32, make_uint4(U[0], U[1], U[2], U[3])); sts128(a + 0 * 32 + 16, make_uint4(U[4], U[5], U[6], U[7]));
                            sts128(a + 1 * 32, make_uint4(V[0], V[1], V[2], V[3])); sts128(a + 1 * 32 + 16, make_uint4(V[4], V[5], V[6], V[7]));
                            sts128(a + 2 * 32, make_uint4(X[0], X[1], X[2], X[3])); sts128(a + 2 * 32 + 16, make_uint4(X[4], X[5], X[6], X[7]));
                            sts128(a + 3 * 32, make_uint4(Y[0], Y[1], Y[2], Y[3])); sts128(a + 3 * 32 + 16, make_uint4(Y[4], Y[5], Y[6], Y[7]));
                            sts128(a + 4 * 32, make_uint4(X2[0], X2[1], X2[2], X2[3])); sts128(a + 4 * 32 + 16, make_uint4(X2[4], X2[5], X2[6], X2[7]));
                            sts128(a + 5 * 32, make_uint4(Y2[0], Y2[1], Y2[2], Y2[3])); sts128(a + 5 * 32 + 16, make_uint4(Y2[4], Y2[5], Y2[6], Y2[7]));
                            sts128(a + 6 * 32, make_uint4(S[0], S[1], S[2], S[3])); sts128(a + 6 * 32 + 16, make_uint4(S[4], S[5], S[6], S[7]));
                            sts128(a + 7 * 32, make_uint4(Hr[0], Hr[1], Hr[2], Hr[3])); sts128(a + 7 * 32 + 16, make_uint4(Hr[4], Hr[5], Hr[6], Hr[7]));
                            sts128(a + 8 * 32, make_uint4((uint32_t)Hb, tw, qw, 0u));
                            has_vec = false; Vt += NT;
                        }
                    }
                    // lane 15 of every main warp's last vector, for the next antidiagonal
                    if (NWM > 1 && lane == 31) {
                        const uint32_t ea = sb + OFF_EDGE + (uint32_t)(par * NWM + warp) * 32u;
                        sts128(ea, make_uint4(X[7], V[7], DUAL ? X2[7] : 0u, qw));
                    }
                } else {
                    // ================= the edge warp: one cell per lane =================
                    const int Vg = grp == 0 ? st_ : en_;
                    const int t = (Vg << 4) + c16;
                    const bool vact = grp == 0 || two;
                    const int Len = (two ? 16 : 0) + (en0 & 15);          // the lane of column en0
                    // a vector that became the lowest one at the end of r-1 arrives from its main thread
                    if (load_lo) {
                        if (grp == 0) {
                            const uint32_t a = sb + OFF_LOST + 4u * (uint32_t)(c16 & 7) + 2u * (uint32_t)(c16 >> 3);
                            eu = lds16(a + 0 * 32); ev = lds16(a + 1 * 32); ex = lds16(a + 2 * 32); ey = lds16(a + 3 * 32);
                            ex2 = lds16(a + 4 * 32); ey2 = lds16(a + 5 * 32); es = lds16(a + 6 * 32);
                            const uint4 e = lds128(sb + OFF_LOST + 8 * 32);
                            eH = (int32_t)e.x + sext16(lds16(a + 7 * 32));
                            etb = (e.y >> (2 * c16)) & 3u; eqb = (e.z >> (2 * c16)) & 3u;
                        }
                        load_lo = false;
                    }
                    // ---- query bases move up one lane (lane c faces query[r - t]); lane 0 reads the new one from memory (prefetched),
                    // lane 16 takes it from lane 15 of the vector below, wherever that one lives
                    const uint32_t q15 = __shfl_sync(FULL, eqb, 15), q31 = __shfl_sync(FULL, eqb, 31);      // before the shift
                    uint32_t qin = __shfl_up_sync(FULL, eqb, 1), qoin = __shfl_up_sync(FULL, qb_ov, 1);
                    uint32_t xt1 = __shfl_up_sync(FULL, ex, 1), vt1 = __shfl_up_sync(FULL, ev, 1), x2t1 = DUAL ? __shfl_up_sync(FULL, ex2, 1) : 0;
                    int32_t hl = __shfl_up_sync(FULL, eH, 1);
                    if (lane == 16) {
                        qoin = two ? q31 : q15;
                        if (two && en_ != st_ + 1) {        // the vector below the highest one is not in group 0
                            if (nb_saved) { xt1 = nbx; vt1 = nbv; x2t1 = nbx2; qin = nbq; hl = nbh; }
                            else {
                                const uint32_t ea = sb + OFF_MHI + (uint32_t)ppar * 32u;
                                const uint4 e = lds128(ea);
                                xt1 = e.x >> 16; vt1 = e.y >> 16; x2t1 = e.z >> 16; qin = e.w >> 30; hl = (int32_t)lds32(ea + 16u);
                            }
                        }
                    }
                    nb_saved = false;
                    const bool inl = st > 0 && st - 1 >= (st0p & ~15) && st - 1 <= (en0p | 15);
                    // carries of the lowest vector (:118-122); x1 / v1 are needed by lanes 1..3 too (single-affine: _mm_cvtsi32_si128(int8_t)
                    // sign-extends a negative carry into them, :146-147)
                    {
                        uint32_t vfirst = 0;
                        if (__builtin_expect(st == 0, 0)) {
                            if (DUAL) vfirst = hi8(r == 0 ? -K.qe : r < sc.long_thres ? -sc.e : r == sc.long_thres ? sc.long_diff : -sc.e2);
                            else vfirst = hi8(r ? sc.q : 0);
                        }
                        const uint32_t x1 = inl ? cx : (K.gX & 0xffffu);
                        const uint32_t v1 = inl ? cv : st > 0 ? (K.gU & 0xffffu) : vfirst;
                        const uint32_t x21 = inl ? cx2 : (K.gX2 & 0xffffu);
                        if (lane == 0) { xt1 = x1; vt1 = v1; x2t1 = x21; qin = q_lo & 3u; hl = ch; }
                        if (!DUAL && lane >= 1 && lane <= 3) {
                            if (x1 & 0x8000u) xt1 = 0xff00u | KC::cE;
                            if (v1 & 0x8000u) vt1 = 0xff00u;
                        }
                    }
                    if (vact) eqb = qin;
                    if (grp == 1) qb_ov = qoin;
                    // ---- profile stores (:126-140): whole 16-lane stores from st0, so the last one overhangs en0 (possibly into the vector above)
                    const int store_end = st0 + ((en0 - st0) >> 4) * 16 + 15;
                    if (vact && t >= st0 && t <= store_end) es = ((etb == eqb ? K.sMch : K.sMis) & 0xffffu);
                    if (grp == 1 && ((en_ + 1) << 4) + c16 <= store_end) s_ov = ((tb_ov == qb_ov ? K.sMch : K.sMis) & 0xffffu);
                    // ---- first row (:123): only while r <= w
                    if (__builtin_expect(en >= r, 0)) {
                        if (vact && t == r) {
                            uint32_t e1;
                            if (DUAL) e1 = hi8(r == 0 ? -K.qe : r < sc.long_thres ? -sc.e : r == sc.long_thres ? sc.long_diff : -sc.e2);
                            else e1 = hi8(r ? sc.q : 0);
                            eu = e1; ey = K.gY & 0xffffu;
                            if (DUAL) ey2 = K.gY2 & 0xffffu;
                        }
                    }
                    // ---- H[en0-1] as the reference's H[] holds it (:231)
                    {
                        const int32_t hk_new = (r > 0 && en0 > 0 && en0 - 1 >= st0p) ? hl : hk;
                        hk = __shfl_sync(FULL, hk_new, Len);
                    }
                    // ---- the cell, its traceback byte, H
                    if (vact) {
                        const uint32_t tbb = ew_cell<DUAL>(eu, ev, ex, ey, ex2, ey2, es, xt1, vt1, x2t1, K, ALL1);
                        tb_page[(int64_t)tb_rip * T.pitch + (t - st)] = (uint8_t)tbb;
                        const int vv = DUAL ? (int)(int8_t)(ev >> 8) : (int)((ev >> 8) & 0xffu) - K.bias;
                        eH += vv;
                    }
                    {
                        // lane en0 takes the value derived from its left neighbour (:231, and :262 for the very first cell)
                        const uint32_t u16 = r == 0 ? ev : eu;
                        const int un = DUAL ? (int)(int8_t)(u16 >> 8) : (int)((u16 >> 8) & 0xffu);
                        if (lane == Len && (r == 0 || en0 > 0)) eH = (r == 0 ? -K.r0_bias : hk - K.bias) + un;
                    }
                    const bool inb = vact && t >= st0 && t <= en0;
                    habs = inb ? eH : INT32_MIN;
                    tp = inb ? t : -1; Hp = eH;
                    if (lane == Len && en0 == tlen - 1) sts32(sb + OFF_HEN0 + 4u * (uint32_t)s3, (uint32_t)eH);
                    if (inb && t == st0 && r - st0 == qlen - 1) sts32(sb + OFF_HST0 + 4u * (uint32_t)s3, (uint32_t)eH);
                    // lane 15 of the lowest vector, for the main thread of the vector above it
                    if (lane == 15) {
                        const uint32_t ea = sb + OFF_EWLO + (uint32_t)par * 32u;
                        sts128(ea, make_uint4(ex << 16, ev << 16, DUAL ? ex2 << 16 : 0u, eqb << 30));
                    }
                    // ================= vectors change places for antidiagonal r+1 =================
                    if (__builtin_expect(nvalid && (st_n != st_ || en_n != en_), 0)) {
                        const bool hi_adv = en_n > en_, lo_adv = st_n > st_;
                        // (1) the highest vector enters the interior: packed form to shared memory for its main thread
                        if (hi_adv && two && en_ > st_n) {
                            const int32_t hb16 = __shfl_sync(FULL, eH, 16);
                            const uint32_t twp = __reduce_or_sync(FULL, grp == 1 ? etb << (2 * c16) : 0u), qwp = __reduce_or_sync(FULL, grp == 1 ? eqb << (2 * c16) : 0u);
                            if (grp == 1) {
                                const uint32_t a = sb + OFF_HIST + 4u * (uint32_t)(c16 & 7) + 2u * (uint32_t)(c16 >> 3);
                                sts16(a + 0 * 32, eu); sts16(a + 1 * 32, ev); sts16(a + 2 * 32, ex); sts16(a + 3 * 32, ey);
                                sts16(a + 4 * 32, ex2); sts16(a + 5 * 32, ey2); sts16(a + 6 * 32, es); sts16(a + 7 * 32, (uint32_t)(eH - hb16));
                                if (lane == 16) sts128(sb + OFF_HIST + 8 * 32, make_uint4((uint32_t)hb16, twp, qwp, 0u));
                            }
                            // its lane 15 is what lane 16 (the new highest vector's first cell) needs at r+1
                            const uint32_t a0 = __shfl_sync(FULL, ex, 31), a1 = __shfl_sync(FULL, ev, 31), a2 = __shfl_sync(FULL, ex2, 31), a3 = __shfl_sync(FULL, eqb, 31);
                            const int32_t a4 = __shfl_sync(FULL, eH, 31);
                            nbx = a0; nbv = a1; nbx2 = a2; nbq = a3; nbh = a4; nb_saved = true;
                        }
                        // (2) the lowest vector leaves the band: its lane 15 is the carry of the next one (:118-122)
                        if (lo_adv) {
                            const uint32_t c0 = __shfl_sync(FULL, ex, 15), c1 = __shfl_sync(FULL, ev, 15), c2 = __shfl_sync(FULL, ex2, 15);
                            ch = __shfl_sync(FULL, eH, 15);
                            // the new lowest vector: group 1's (the band has shrunk to it), the vector above the band (band of one vector
                            // moving up), or an interior one from its main thread
                            const uint32_t m0 = __shfl_down_sync(FULL, eu, 16), m1 = __shfl_down_sync(FULL, ev, 16), m2 = __shfl_down_sync(FULL, ex, 16), m3 = __shfl_down_sync(FULL, ey, 16);
                            const uint32_t m4 = __shfl_down_sync(FULL, ex2, 16), m5 = __shfl_down_sync(FULL, ey2, 16), m6 = __shfl_down_sync(FULL, es, 16);
                            const uint32_t m7 = __shfl_down_sync(FULL, etb, 16), m8 = __shfl_down_sync(FULL, eqb, 16);
                            const int32_t m9 = __shfl_down_sync(FULL, eH, 16);
                            const uint32_t o6 = __shfl_down_sync(FULL, s_ov, 16), o7 = __shfl_down_sync(FULL, tb_ov, 16), o8 = __shfl_down_sync(FULL, qb_ov, 16);
                            if (grp == 0) {
                                if (two && st_n == en_) { eu = m0; ev = m1; ex = m2; ey = m3; ex2 = m4; ey2 = m5; es = m6; etb = m7; eqb = m8; eH = m9; }
                                else if (st_n == en_ + 1) {
                                    eu = K.gU & 0xffffu; ev = K.gU & 0xffffu; ex = K.gX & 0xffffu; ey = K.gY & 0xffffu; ex2 = K.gX2 & 0xffffu; ey2 = K.gY2 & 0xffffu;
                                    es = o6; etb = o7; eqb = o8; eH = 0;
                                } else load_lo = true;
                            }
                            load_lo = __shfl_sync(FULL, load_lo ? 1 : 0, 0) != 0;
                            cx = c0; cv = c1; cx2 = c2;
                        }
                        // (3) the vector above the band becomes the highest one; a new one above it
                        if (hi_adv) {
                            if (grp == 1) {
                                if (en_n > st_n) {
                                    eu = K.gU & 0xffffu; ev = K.gU & 0xffffu; ex = K.gX & 0xffffu; ey = K.gY & 0xffffu; ex2 = K.gX2 & 0xffffu; ey2 = K.gY2 & 0xffffu;
                                    es = s_ov; etb = tb_ov; eqb = qb_ov; eH = 0;
                                }
                                const int tv = ((en_n + 1) << 4) + c16, j = r - tv;
                                s_ov = K.sInit & 0xffffu;
                                tb_ov = tv < tlen ? (uint32_t)target[tv] & 3u : 0u;
                                qb_ov = (j >= 0 && j < qlen) ? (uint32_t)query[j] & 3u : 0u;
                            }
                        }
                    }
                    // the query base lane 0 needs at r+1
                    if (lane == 0 && nvalid) { const int j = r + 1 - (st_n << 4); q_lo = (j >= 0 && j < qlen) ? (uint32_t)query[j] : 0u; }
                }
                habs_p = habs; st0p = st0; en0p = en0; st_p = st_; en_p = en_;
                grad_p = nvalid && en_n > en_ && two && en_ > st_n; grad_v = en_;
                {
                    const int32_t wmax = __reduce_max_sync(FULL, habs);
                    if (lane == 0) {
                        const uint32_t a = sb + OFF_MX + 4u * (uint32_t)(s3 * NSLOT + warp);
                        if (!(warp == 0 && dropped)) sts32(a, (uint32_t)wmax);
                    }
                }
                if (__builtin_expect(++tb_rip == T.rows_per_page, 0)) {
                    tb_rip = 0; ++tb_pg;
                    if (tb_pg < T.tb_pages) tb_page = C.pool.base + (int64_t)table[tb_pg] * C.pool.page_bytes;
                    if (tid == 0 && tb_pg + 1 < T.tb_pages) {
                        const int held = (int)lds32(sb + OFF_HELD);
                        if (held < tb_pg + 2) sts32(sb + OFF_HELD, (uint32_t)pool_lazy_grow(C.pool, C.slot_base + (int)blockIdx.x, table, held, tb_pg + 2));
                    }
                }
            }
            st0 = st0n; en0 = en0n;
            if (r + 1 >= n_diag && valid) { st0 = 1; en0 = 0; }       // behind the last antidiagonal
            // ---- (D) the one barrier of the antidiagonal
            __syncthreads();
            M2 = M1; nt2 = nt1;
            s3 = s3 == 2 ? 0 : s3 + 1;
        }
        if (!dropped && stop_r < n_diag) ez.zdropped = 1;      // band exhausted (:111-114)
        __syncthreads();
        if (warp == 0) finish_task(C, T, table, ez, cells, true);
        __syncthreads();
        if (tid == 0) {
            const bool lazy = C.pool.lazy && C.pool.lazy_min_pages > 0 && T.tb_pages >= C.pool.lazy_min_pages;
            const int held = (int)lds32(sb + OFF_HELD);
            if (!lazy && pool_may_keep(C.pool, held)) kept = held;      // table[0 .. held) stays with this CTA for its next task
            else pool_free(C.pool, held, table, lazy ? C.slot_base + (int)blockIdx.x : -1);
            if (C.timeline) C.timeline[2 * T.orig + 1] = global_ns();
        }
    }
}

// ---------------------------------------------------------------------------
// host side

// main warps a task needs here: one thread per vector that can be strictly inside the band at the same time, and vector V lives in
// thread V mod NT, so NT >= (vectors of the band) - 1
inline int ew_warps_needed(const DevTask& t) { return (t.pitch / 16 - 1 + 31) / 32; }

// tasks the edge-warp kernel takes (the others stay with fsv_fill_dpx_kernel)
inline bool ew_supports(const DevScoring& sc, const DevTask& t, bool has_wild)
{
    if (!dpx_supports(sc, t, has_wild) || has_wild) return false;
    if (t.flag & (FSV_EZ_RIGHT | FSV_EZ_SCORE_ONLY | FSV_EZ_APPROX_MAX)) return false;
    if (t.tb_pages <= 0 || t.w < 32 || t.qlen < 64 || t.tlen < 64) return false;
    return dpx_class_of(ew_warps_needed(t)) != 0;
}

template <bool DUAL, int NWM>
inline int ew_grid_one(int sm_count, int n_tasks)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fsv_fill_ew_kernel<DUAL, NWM, false>, (NWM + 1) * 32, 0) != cudaSuccess) { cudaGetLastError(); per_sm = 1; }
    if (per_sm < 1) per_sm = 1;
    if (getenv("FSV_TRACE")) fprintf(stderr, "[fsv] edge-warp kernel, %d main warps: %d CTAs per SM\n", NWM, per_sm);
    return std::max(1, std::min(n_tasks, sm_count * per_sm));
}
template <bool DUAL, int NWM>
inline int ew_launch_one(cudaStream_t stream, int grid, bool excl, const DpxParams& P, std::string* err)
{
    cudaError_t e = cudaSuccess;
    DpxParams PK = P;
    PK.K = DpxConst<DUAL, false>(P.C.sc);
    if (excl) {
        auto kern = fsv_fill_ew_kernel<DUAL, NWM, true>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DPX_EXCL_SMEM - 8192);
        if (e == cudaSuccess) kern<<<grid, (NWM + 1) * 32, DPX_EXCL_SMEM - 8192, stream>>>(PK);
    } else {
        fsv_fill_ew_kernel<DUAL, NWM, false><<<grid, (NWM + 1) * 32, 0, stream>>>(PK);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { if (err) *err = cudaGetErrorString(e); cudaGetLastError(); return FSV_ERR_CUDA; }
    return FSV_OK;
}
template <bool DUAL>
inline int ew_grid_nw(int sm_count, int nw, int n_tasks)
{
    switch (nw) {
        case 1: return ew_grid_one<DUAL, 1>(sm_count, n_tasks);
        case 2: return ew_grid_one<DUAL, 2>(sm_count, n_tasks);
        case 4: return ew_grid_one<DUAL, 4>(sm_count, n_tasks);
        case 6: return ew_grid_one<DUAL, 6>(sm_count, n_tasks);
        case 8: return ew_grid_one<DUAL, 8>(sm_count, n_tasks);
    }
    return 1;
}
template <bool DUAL>
inline int ew_launch_nw(cudaStream_t stream, int nw, int grid, bool excl, const DpxParams& P, std::string* err)
{
    switch (nw) {
        case 1: return ew_launch_one<DUAL, 1>(stream, grid, excl, P, err);
        case 2: return ew_launch_one<DUAL, 2>(stream, grid, excl, P, err);
        case 4: return ew_launch_one<DUAL, 4>(stream, grid, excl, P, err);
        case 6: return ew_launch_one<DUAL, 6>(stream, grid, excl, P, err);
        case 8: return ew_launch_one<DUAL, 8>(stream, grid, excl, P, err);
    }
    return FSV_ERR_INVALID;
}
// compiled in the (DUAL, traceback) translation units of fsv_dpx_variant.cu
int ew_grid_0(int sm_count, int nw, int n_tasks);
int ew_grid_1(int sm_count, int nw, int n_tasks);
int ew_launch_0(cudaStream_t stream, int nw, int grid, bool excl, const DpxParams& P, std::string* err);
int ew_launch_1(cudaStream_t stream, int nw, int grid, bool excl, const DpxParams& P, std::string* err);

}  // namespace fsv
