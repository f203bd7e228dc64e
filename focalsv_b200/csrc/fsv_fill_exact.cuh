// fsv_fill_exact.cuh — the GENERAL fill kernel: one CTA per task, one thread per
// int8 lane, every lane computed with the reference's exact int8 semantics.
//
// This kernel takes any task the ABI accepts (any band, any length, wildcard
// bases, KSW_EZ_GENERIC_SC, approximate-max mode, either gap model).  The
// register-resident DPX kernel in fsv_fill_dpx.cuh handles the common shapes
// faster; tasks it does not cover are routed here.  Both are CUDA; there is no
// CPU path.
//
// Follows software/hifiasm-0.16.1/ksw2_extz2_sse.c:101-289 lane by lane; the
// dual-affine branch follows the restatement documented in DESIGN.md §3.
#pragma once
#include "fsv_backtrack.cuh"
#include "fsv_common.cuh"

namespace fsv {

constexpr int EXACT_THREADS = 256;
constexpr int EXACT_WARPS = EXACT_THREADS / 32;
constexpr int EXACT_NARR = 10;  // u v[2] x[2] y x2[2] y2 s

struct FillParams {
    RunCtx C;
    TaskQueue Q;
    uint8_t* ws;             // per-CTA global window for tasks whose band exceeds shared memory
    int64_t ws_lanes;        // lanes per array in that window (power of two), 0 = none
    int32_t smem_lanes;      // lanes per array of the shared-memory window (power of two)
};

__device__ __forceinline__ int s8(int x) { return (int)(int8_t)(uint8_t)x; }

// the circular lane window of one task
struct LaneWindow {
    uint8_t *u, *v[2], *x[2], *y, *x2[2], *y2, *s;
    int32_t* H;
    uint32_t mask;
    __device__ __forceinline__ uint32_t at(int t) const { return (uint32_t)t & mask; }
};

template <bool DUAL>
__device__ __forceinline__ void init_lanes(const LaneWindow& W, const DevScoring& sc, int lo, int hi)
{   // arrays start zeroed (kcalloc, ksw2_extz2_sse.c:84) or at -(q+e) / -(q2+e2) (dual-affine memset)
    const uint8_t g1 = DUAL ? (uint8_t)(-sc.q - sc.e) : 0, g2 = DUAL ? (uint8_t)(-sc.q2 - sc.e2) : 0;
    for (int t = lo + (int)threadIdx.x; t <= hi; t += EXACT_THREADS) {
        uint32_t k = W.at(t);
        W.u[k] = g1; W.v[0][k] = g1; W.v[1][k] = g1; W.x[0][k] = g1; W.x[1][k] = g1; W.y[k] = g1;
        if (DUAL) { W.x2[0][k] = g2; W.x2[1][k] = g2; W.y2[k] = g2; }
        W.s[k] = 0;
        W.H[k] = FSV_NEG_INF;
    }
}

template <bool DUAL>
__global__ void __launch_bounds__(EXACT_THREADS) fsv_fill_exact_kernel(const FillParams P)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ int32_t sh_task;
    __shared__ int32_t sh_part_h[2][EXACT_WARPS];
    __shared__ uint32_t sh_part_k[2][EXACT_WARPS];
    const RunCtx& C = P.C;
    const DevScoring& sc = C.sc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int32_t* table = C.page_tables + (int64_t)blockIdx.x * C.max_pages_per_task;
    int pending = -1;

    for (;;) {
        __syncthreads();
        if (tid == 0) { int held_, kept_ = 0; sh_task = next_task(C, P.Q, table, pending, false, held_, kept_); }     // always up front, nothing kept
        __syncthreads();
        const int ti = sh_task;
        if (ti < 0) return;
        const DevTask T = C.tasks[ti];
        if (tid == 0 && C.timeline) C.timeline[2 * T.orig] = global_ns();
        EzState ez; ez.reset();
        int64_t cells = 0;
        if (T.kind == 0) {           // ksw2's silent returns (ksw2_extz2_sse.c:57,82)
            if (tid == 0) finish_reset_task(C, T);
            continue;
        }
        const int qlen = T.qlen, tlen = T.tlen, w = T.w, flag = T.flag;
        const bool with_cigar = !(flag & FSV_EZ_SCORE_ONLY), right = (flag & FSV_EZ_RIGHT) != 0;
        const bool approx = (flag & FSV_EZ_APPROX_MAX) != 0, generic = (flag & FSV_EZ_GENERIC_SC) != 0;
        const int L = (tlen + 15) / 16 * 16;
        const uint8_t* query = C.qarena + T.q_off;
        const uint8_t* target = C.tarena + T.t_off;

        // lane window: shared memory when the band fits, else this CTA's global slice
        LaneWindow W;
        {
            const int need = T.pitch + 96;      // live lanes: [st-1, en+80]
            uint8_t* base; int64_t lanes;
            if (need <= P.smem_lanes) { base = smem_raw; lanes = P.smem_lanes; }
            else { lanes = P.ws_lanes; base = P.ws + (int64_t)blockIdx.x * lanes * (EXACT_NARR + 4); }
            W.mask = (uint32_t)(lanes - 1);
            W.H = (int32_t*)base; base += lanes * 4;
            W.u = base; W.v[0] = base + lanes; W.v[1] = base + 2 * lanes; W.x[0] = base + 3 * lanes;
            W.x[1] = base + 4 * lanes; W.y = base + 5 * lanes; W.x2[0] = base + 6 * lanes;
            W.x2[1] = base + 7 * lanes; W.y2 = base + 8 * lanes; W.s = base + 9 * lanes;
        }
        const int qe = sc.q + sc.e;
        const int bias = DUAL ? 0 : qe, r0_bias = DUAL ? qe : 2 * qe;
        const uint8_t g1 = DUAL ? (uint8_t)(-sc.q - sc.e) : 0, g2 = DUAL ? (uint8_t)(-sc.q2 - sc.e2) : 0;
        const int m1 = sc.m - 1;
        int init_hi = -1, last_st = -1, last_en = -1, par = 0;
        int32_t H0 = 0, last_H0_t = 0;   // approximate-max cursor (:270-286)
        const int n_diag = qlen + tlen - 1;

        for (int r = 0; r < n_diag; ++r) {
            int st0, en0;
            band_limits(r, qlen, tlen, w, st0, en0);
            if (st0 > en0) { ez.zdropped = 1; break; }            // :111-114
            const int st = round_st(st0), en = round_en(en0);
            cells += en0 - st0 + 1;
            if (en + 16 > init_hi) {                              // expose fresh lanes ahead of the band
                int hi = en + 16 + 64;
                init_lanes<DUAL>(W, sc, init_hi + 1, hi);
                init_hi = hi;
                __syncthreads();
            }
            const int store_end = st0 + ((en0 - st0) / 16) * 16 + 15;  // last lane the profile stores touch (:126-140)
            uint8_t* tbrow = with_cigar ? tb_row(C.pool, table, T.rows_per_page, T.pitch, r) : nullptr;
            const uint8_t* xin = W.x[par]; const uint8_t* vin = W.v[par]; const uint8_t* x2in = W.x2[par];
            uint8_t* xout = W.x[par ^ 1]; uint8_t* vout = W.v[par ^ 1]; uint8_t* x2out = W.x2[par ^ 1];
            // carries into the first lane (:118-122)
            int x1c, v1c, x21c = 0;
            if (st > 0) {
                if (st - 1 >= last_st && st - 1 <= last_en) {
                    x1c = xin[W.at(st - 1)]; v1c = vin[W.at(st - 1)];
                    if (DUAL) x21c = x2in[W.at(st - 1)];
                } else { x1c = g1; v1c = g1; x21c = g2; }
            } else {
                x1c = g1; x21c = g2;
                if (DUAL) v1c = (uint8_t)(r == 0 ? -sc.q - sc.e : r < sc.long_thres ? -sc.e : r == sc.long_thres ? sc.long_diff : -sc.e2);
                else v1c = r ? (uint8_t)sc.q : 0;
            }
            int edge;   // first-row value of u (:123)
            if (DUAL) edge = (uint8_t)(r == 0 ? -sc.q - sc.e : r < sc.long_thres ? -sc.e : r == sc.long_thres ? sc.long_diff : -sc.e2);
            else edge = r ? (uint8_t)sc.q : 0;
            const bool top_edge = en >= r;
            int32_t H_left = 0;   // H[en0-1] of the previous antidiagonal, read before anyone updates it
            if (!approx && r > 0 && en0 > 0 && tid == (en0 - st0) % EXACT_THREADS) H_left = W.H[W.at(en0 - 1)];
            const uint8_t* qrr = query;   // qrr[t] = query[r - t] (reversed query of :98,104)

            const int lane_hi = max(en, min(store_end, L - 1));   // DP lanes [st,en] plus the profile overhang
            for (int t = st + tid; t <= lane_hi; t += EXACT_THREADS) {
                const uint32_t k = W.at(t);
                // ---- score profile (:125-144)
                int sv;
                if (t >= st0 && t <= store_end && (!generic || t <= en0)) {
                    int sq = t < tlen ? target[t] : 0;             // zero padding past the target (kcalloc)
                    int j = r - t;
                    int sr = (j >= 0 && j < qlen) ? qrr[j] : 0;
                    if (generic) sv = (uint8_t)sc.mat[sq * sc.m + sr];
                    else {
                        sv = sq == sr ? sc.sc_mch : sc.sc_mis;
                        if (sq == m1 || sr == m1) sv = sc.sc_N;
                    }
                    W.s[k] = (uint8_t)sv;
                } else sv = W.s[k];
                if (t > en) continue;                              // overhang lanes only refresh s
                // ---- operands
                int xt1, vt1, x2t1 = 0;
                if (t == st) { xt1 = x1c; vt1 = v1c; x2t1 = x21c; }
                else {
                    const uint32_t k1 = W.at(t - 1);
                    xt1 = xin[k1]; vt1 = vin[k1]; if (DUAL) x2t1 = x2in[k1];
                    if (!DUAL && t <= st + 3) {   // _mm_cvtsi32_si128(int8_t) sign-extends into lanes 1..3 (:146-147)
                        if (s8(x1c) < 0) xt1 = 0xff;
                        if (s8(v1c) < 0) vt1 = 0xff;
                    }
                }
                int ut, yt, y2t = 0;
                if (top_edge && t == r) { ut = edge; yt = g1; y2t = g2; }
                else { ut = W.u[k]; yt = W.y[k]; if (DUAL) y2t = W.y2[k]; }
                int d = 0, un, vn, xn, yn, x2n = 0, y2n = 0;
                if (!DUAL) {
                    int z = (sv + 2 * qe) & 255, a = (xt1 + vt1) & 255, b = (yt + ut) & 255;
                    if (with_cigar && !right) {
                        d = s8(a) > s8(z) ? 1 : 0;
                        z = s8(z) > s8(a) ? z : a;
                        d = s8(b) > s8(z) ? 2 : d;
                    } else if (with_cigar) {
                        d = s8(z) > s8(a) ? 0 : 1;
                        z = s8(z) > s8(a) ? z : a;
                        d = s8(z) > s8(b) ? d : 2;
                    } else z = s8(z) > s8(a) ? z : a;
                    z = z > b ? z : b;                              // unsigned (:41)
                    z = z < sc.max_sc_clamp ? z : sc.max_sc_clamp;  // unsigned (:42)
                    un = (z - vt1) & 255; vn = (z - ut) & 255;
                    z = (z - sc.q) & 255; a = (a - z) & 255; b = (b - z) & 255;
                    if (!with_cigar || !right) {
                        int ca = s8(a) > 0, cb = s8(b) > 0;
                        xn = ca ? a : 0; yn = cb ? b : 0;
                        d |= (ca ? 0x08 : 0) | (cb ? 0x10 : 0);
                    } else {
                        int na = 0 > s8(a), nb = 0 > s8(b);
                        xn = na ? 0 : a; yn = nb ? 0 : b;
                        d |= (na ? 0 : 0x08) | (nb ? 0 : 0x10);
                    }
                } else {
                    int z = sv, a = (xt1 + vt1) & 255, b = (yt + ut) & 255, a2 = (x2t1 + vt1) & 255, b2 = (y2t + ut) & 255;
                    if (with_cigar && !right) {
                        d = s8(a) > s8(z) ? 1 : 0;  z = s8(z) > s8(a) ? z : a;
                        d = s8(b) > s8(z) ? 2 : d;  z = s8(z) > s8(b) ? z : b;
                        d = s8(a2) > s8(z) ? 3 : d; z = s8(z) > s8(a2) ? z : a2;
                        d = s8(b2) > s8(z) ? 4 : d; z = s8(z) > s8(b2) ? z : b2;
                    } else if (with_cigar) {
                        d = s8(z) > s8(a) ? 0 : 1;  z = s8(z) > s8(a) ? z : a;
                        d = s8(z) > s8(b) ? d : 2;  z = s8(z) > s8(b) ? z : b;
                        d = s8(z) > s8(a2) ? d : 3; z = s8(z) > s8(a2) ? z : a2;
                        d = s8(z) > s8(b2) ? d : 4; z = s8(z) > s8(b2) ? z : b2;
                    } else {
                        z = s8(z) > s8(a) ? z : a; z = s8(z) > s8(b) ? z : b;
                        z = s8(z) > s8(a2) ? z : a2; z = s8(z) > s8(b2) ? z : b2;
                    }
                    z = s8(z) < s8(sc.max_sc_clamp) ? z : sc.max_sc_clamp;
                    un = (z - vt1) & 255; vn = (z - ut) & 255;
                    int tmp = (z - sc.q) & 255; a = (a - tmp) & 255; b = (b - tmp) & 255;
                    tmp = (z - sc.q2) & 255; a2 = (a2 - tmp) & 255; b2 = (b2 - tmp) & 255;
                    int ca, cb, ca2, cb2;
                    if (!with_cigar || !right) { ca = s8(a) > 0; cb = s8(b) > 0; ca2 = s8(a2) > 0; cb2 = s8(b2) > 0; }
                    else { ca = !(0 > s8(a)); cb = !(0 > s8(b)); ca2 = !(0 > s8(a2)); cb2 = !(0 > s8(b2)); }
                    xn = ((ca ? a : 0) - qe) & 255; yn = ((cb ? b : 0) - qe) & 255;
                    x2n = ((ca2 ? a2 : 0) - (sc.q2 + sc.e2)) & 255; y2n = ((cb2 ? b2 : 0) - (sc.q2 + sc.e2)) & 255;
                    d |= (ca ? 0x08 : 0) | (cb ? 0x10 : 0) | (ca2 ? 0x20 : 0) | (cb2 ? 0x40 : 0);
                }
                W.u[k] = (uint8_t)un; vout[k] = (uint8_t)vn; xout[k] = (uint8_t)xn; W.y[k] = (uint8_t)yn;
                if (DUAL) { x2out[k] = (uint8_t)x2n; W.y2[k] = (uint8_t)y2n; }
                if (with_cigar) tbrow[t - st] = (uint8_t)d;
            }
            __syncthreads();   // A: every lane of this antidiagonal is in the window

            int dropped = 0;
            if (!approx) {     // exact max with the 32-bit H array (:224-269)
                const int en1 = st0 + (en0 - st0) / 4 * 4;
                int32_t best_h = INT32_MIN; uint32_t best_k = 0xffffffffu;
                for (int t = st0 + tid; t <= en0; t += EXACT_THREADS) {
                    const uint32_t k = W.at(t);
                    int32_t h; uint32_t key;
                    if (r == 0) { h = (DUAL ? s8(vout[k]) : (int)vout[k]) - r0_bias; key = 1u + (4u << 26) + (uint32_t)t; }
                    else if (t < en0) {
                        h = W.H[k] + (DUAL ? s8(vout[k]) : (int)vout[k]) - bias;
                        key = t < en1 ? 1u + ((uint32_t)((t - st0) & 3) << 26) + (uint32_t)t : 1u + (4u << 26) + (uint32_t)t;
                    } else {
                        h = en0 > 0 ? H_left + (DUAL ? s8(W.u[k]) : (int)W.u[k]) - bias
                                    : W.H[k] + (DUAL ? s8(vout[k]) : (int)vout[k]) - bias;
                        key = 0;   // H[en0] seeds the scan, so it wins every tie (:231-232)
                    }
                    W.H[k] = h;
                    if (h > best_h || (h == best_h && key < best_k)) { best_h = h; best_k = key; }
                }
                const int32_t wm = __reduce_max_sync(0xffffffffu, best_h);
                const uint32_t wk = __reduce_min_sync(0xffffffffu, best_h == wm ? best_k : 0xffffffffu);
                if (lane == 0) { sh_part_h[r & 1][warp] = wm; sh_part_k[r & 1][warp] = wk; }
                __syncthreads();   // B
                int32_t max_H = INT32_MIN; uint32_t mk = 0xffffffffu;
#pragma unroll
                for (int i = 0; i < EXACT_WARPS; ++i) {
                    int32_t h = sh_part_h[r & 1][i]; uint32_t k = sh_part_k[r & 1][i];
                    if (h > max_H || (h == max_H && k < mk)) { max_H = h; mk = k; }
                }
                const int max_t = mk == 0 ? en0 : (int)((mk - 1u) & ((1u << 26) - 1u));
                const int32_t H_en0 = W.H[W.at(en0)], H_st0 = W.H[W.at(st0)];
                if (en0 == tlen - 1 && H_en0 > ez.mte) { ez.mte = H_en0; ez.mte_q = r - en; }   // rounded en (:263-264)
                if (r - st0 == qlen - 1 && H_st0 > ez.mqe) { ez.mqe = H_st0; ez.mqe_t = st0; }
                dropped = ez.apply_zdrop(max_H, r, max_t, T.zdrop, sc.e_drop);
                if (!dropped && r == n_diag - 1 && en0 == tlen - 1) ez.score = W.H[W.at(tlen - 1)];
            } else {           // approximate max: follow one cell (:270-286)
                if (r > 0) {
                    const bool in0 = last_H0_t >= st0 && last_H0_t <= en0;
                    const bool in1 = last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0;
                    if (in0 && in1) {
                        int d0 = (DUAL ? s8(vout[W.at(last_H0_t)]) : (int)vout[W.at(last_H0_t)]) - bias;
                        int d1 = (DUAL ? s8(W.u[W.at(last_H0_t + 1)]) : (int)W.u[W.at(last_H0_t + 1)]) - bias;
                        if (d0 > d1) H0 += d0; else { H0 += d1; ++last_H0_t; }
                    } else if (in0) {
                        H0 += (DUAL ? s8(vout[W.at(last_H0_t)]) : (int)vout[W.at(last_H0_t)]) - bias;
                    } else {
                        ++last_H0_t;
                        H0 += (DUAL ? s8(W.u[W.at(last_H0_t)]) : (int)W.u[W.at(last_H0_t)]) - bias;
                    }
                    if (!DUAL) { if ((flag & FSV_EZ_APPROX_DROP) && ez.apply_zdrop(H0, r, last_H0_t, T.zdrop, sc.e_drop)) dropped = 1; }
                } else { H0 = (DUAL ? s8(vout[W.at(0)]) : (int)vout[W.at(0)]) - r0_bias; last_H0_t = 0; }
                if (DUAL) { if ((flag & FSV_EZ_APPROX_DROP) && ez.apply_zdrop(H0, r, last_H0_t, T.zdrop, sc.e_drop)) dropped = 1; }
                if (!dropped && r == n_diag - 1 && en0 == tlen - 1) ez.score = H0;
                __syncthreads();   // B: u/v of this antidiagonal may be overwritten from here on
            }
            if (dropped) break;
            last_st = st; last_en = en; par ^= 1;
        }

        __syncthreads();             // every traceback row is written
        if (warp == 0) finish_task(C, T, table, ez, cells, with_cigar);
        __syncthreads();
        if (tid == 0) { pool_free(C.pool, T.tb_pages, table); if (C.timeline) C.timeline[2 * T.orig + 1] = global_ns(); }
    }
}

}  // namespace fsv
