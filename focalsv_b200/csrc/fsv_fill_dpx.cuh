// fsv_fill_dpx.cuh — the FAST fill kernel: register-resident, DPX (16x2) arithmetic.
//
// Shape (B200 / sm_100a):
//   * one thread == one 16-lane vector of the reference's SSE loop
//     (software/hifiasm-0.16.1/ksw2_extz2_sse.c:151,174,200: `for (t = st_; t <= en_; ++t)`),
//     so the reference's rounding of the band to 16-lane vectors (:116) maps onto
//     whole threads and every lane it computes outside [st0,en0] is computed here too;
//   * the vector's state u,v,x,y,(x2,y2),s lives in REGISTERS for the whole time the
//     vector is inside the band (about 2w antidiagonals), as 8 packed 16x2 words per array:
//     word k holds lanes (k, k+8) of the vector, so the "t-1" operand of word k is simply
//     word k-1 (no shifting); only word 0 needs the neighbour thread (one shuffle);
//   * every int8 lane of the reference is kept in the HIGH byte of a 16-bit half: 16-bit
//     wrap-around == int8 wrap-around, 16-bit signed/unsigned compares == int8 compares,
//     and the LOW byte is free to carry a priority code, so that one VIMNMX3.S16x2
//     yields both max(...) and which operand won, with the reference's tie order;
//   * the band slides by one lane every other antidiagonal: vectors are owned
//     round-robin (vector V -> thread V mod NT); a thread whose vector leaves the band
//     re-arms itself NT vectors further right;
//   * per antidiagonal: one 16-byte traceback store per thread (coalesced 512 B per warp),
//     one warp REDUX for the exact max, one CTA barrier pair when a task spans several warps.
// Tasks it does not take (wildcard bases, KSW_EZ_GENERIC_SC / RIGHT / APPROX_MAX, bands
// wider than 8 warps of vectors) go to the general kernel in fsv_fill_exact.cuh.
#pragma once
#include <string>
#include <type_traits>

#include "fsv_backtrack.cuh"
#include "fsv_common.cuh"

namespace fsv {

constexpr int DPX_MAX_WARPS = 8;

// scoring constants in the "int8 in the high byte of a 16-bit half" format.  Built on the HOST per launch and
// passed in the kernel parameters, so that the fill loop reads them as constant-bank operands instead of
// holding (or rebuilding) fifteen registers.
struct DpxK {
    uint32_t gU, gX, gY, gX2, gY2, sInit, sMch, sMis, sAmb, kClamp, kQ, kQ2, kQE, kQE2, kBias;
    uint32_t nClamp, kQ1, kQ21;       // ~kClamp, kQ | 0x0001, kQ2 | 0x0001 per half: see "z is kept complemented" in dpx_cells
    uint32_t nQE, nQE2, nBias;        // -(q+e), -(q2+e2), -bias per half: x - const is x + (-const), one VIADD.16x2
    int qe, qe2, bias, r0_bias;
};

struct DpxParams {
    RunCtx C;
    TaskQueue Q;
    DpxK K;
    int32_t seg_launch;         // SEG launches: which ticket counter (one per launch: a launch's CTAs only ever wait for tasks of their own queue)
};

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
}
__host__ __device__ __forceinline__ uint32_t hi8(int v) { return ((uint32_t)v & 0xffu) << 8; }       // int8 -> high byte of a half
__host__ __device__ __forceinline__ uint32_t both(uint32_t h) { return (h & 0xffffu) * 0x00010001u; } // same half twice
__device__ __forceinline__ int sext16(uint32_t h) { return (int)(int16_t)(uint16_t)h; }

// Lane c of a vector lives in half (c >> 3) of word (c & 7).  Both accessors use static register
// indices only (select trees), so the arrays stay in registers (an `if (k == c) A[k] = ...` chain
// would be turned into a local-memory access by the compiler).
__device__ __forceinline__ void set_cell(uint32_t (&A)[8], int c, uint32_t v16)
{
    const int kk = c & 7;
    const uint32_t keep = (c & 8) ? 0x0000ffffu : 0xffff0000u, val = both(v16) & ~keep;
#pragma unroll
    for (int k = 0; k < 8; ++k) A[k] = (k == kk) ? ((A[k] & keep) | val) : A[k];
}
__device__ __forceinline__ uint32_t get_cell(const uint32_t (&A)[8], int c)
{
    const bool b0 = c & 1, b1 = c & 2, b2 = c & 4;
    const uint32_t s0 = b0 ? A[1] : A[0], s1 = b0 ? A[3] : A[2], s2 = b0 ? A[5] : A[4], s3 = b0 ? A[7] : A[6];
    const uint32_t t0 = b1 ? s1 : s0, t1 = b1 ? s3 : s2;
    const uint32_t w = b2 ? t1 : t0;
    return (c & 8) ? (w >> 16) : (w & 0xffffu);
}
// mask of the halves whose lane index c is in [lo, hi]
__device__ __forceinline__ uint32_t lane_mask(int k, int lo, int hi)
{
    uint32_t m = 0;
    if (k >= lo && k <= hi) m |= 0x0000ffffu;
    if (k + 8 >= lo && k + 8 <= hi) m |= 0xffff0000u;
    return m;
}

// 16-bit lane mask (bit c = lane c) of the lanes in [lo, hi] of a vector
__device__ __forceinline__ uint32_t lane_bits(int lo, int hi)
{
    lo = max(lo, 0); hi = min(hi, 15);
    return hi < lo ? 0u : ((0xffffu << lo) & (0xffffu >> (15 - hi)) & 0xffffu);
}
// packed 16x2 mask of word k (lanes k and k+8) from a 16-bit lane mask
__device__ __forceinline__ uint32_t word_mask(uint32_t bits, int k) { return prmt(bits << (7 - k), 0u, 0x9988u); }

// explicit shared-memory accesses from one base register (the compiler otherwise rebuilds the address of
// every __shared__ array on every use)
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) { asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ void atom_min_shared(uint32_t a, uint32_t v) { asm volatile("red.shared.min.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void red_max_shared(uint32_t a, int32_t v) { asm volatile("red.shared.max.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) { asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(x), "r"(y) : "memory"); }

// resident CTAs per SM the register budget is tuned for
template <int NW> struct DpxOcc { static constexpr int value = NW == 1 ? 12 : NW == 2 ? 6 : NW == 4 ? 3 : 2; };

template <bool DUAL, bool RIGHT = false>
struct DpxConst : DpxK {
    // tie-break codes in the low byte (higher wins).  Left alignment (ksw2_extz2_sse.c:177-181): M > E > F > E~ > F~,
    // the traceback stores the code and the CIGAR walk turns it back into ksw2's d (tb_mode).  Right alignment
    // (KSW_EZ_RIGHT, :203-207): the later state wins a tie, so the code IS ksw2's d.
    static constexpr uint32_t cS = RIGHT ? 0 : DUAL ? 4 : 2, cE = RIGHT ? 1 : DUAL ? 3 : 1, cF = RIGHT ? 2 : DUAL ? 2 : 0,
                              cE2 = RIGHT ? 3 : 1, cF2 = RIGHT ? 4 : 0;
    __host__ __device__ explicit DpxConst(const DevScoring& sc)
    {
        qe = sc.q + sc.e; qe2 = sc.q2 + sc.e2;
        gU = both(DUAL ? hi8(-qe) : 0);                                      // initial u, v (:84 / dual memset)
        gX = both((DUAL ? hi8(-qe) : 0) | cE); gY = both((DUAL ? hi8(-qe) : 0) | cF);
        gX2 = both(hi8(-qe2) | cE2); gY2 = both(hi8(-qe2) | cF2);
        sInit = both((DUAL ? 0 : hi8(2 * qe)) | cS);                         // s[] starts at 0 (kcalloc)
        sMch = both((DUAL ? hi8(sc.sc_mch) : hi8(sc.sc_mch + 2 * qe)) | cS);
        sMis = both((DUAL ? hi8(sc.sc_mis) : hi8(sc.sc_mis + 2 * qe)) | cS);
        sAmb = both((DUAL ? hi8(sc.sc_N) : hi8(sc.sc_N + 2 * qe)) | cS);      // either base is the wildcard m-1 (:68,130)
        kClamp = both(hi8(sc.max_sc_clamp));
        kQ = both(hi8(sc.q)); kQ2 = both(hi8(sc.q2)); kQE = both(hi8(qe)); kQE2 = both(hi8(qe2));
        nClamp = ~kClamp; kQ1 = kQ | 0x00010001u; kQ21 = kQ2 | 0x00010001u;
        nQE = both(hi8(-qe)); nQE2 = both(hi8(-qe2));
        bias = DUAL ? 0 : qe; r0_bias = DUAL ? qe : 2 * qe;
        kBias = both((uint32_t)bias & 0xffffu); nBias = both((uint32_t)(-bias) & 0xffffu);
    }
};

// score profile of the 16 lanes of a vector from the 2-bit packed target / query windows (:125-140)
template <bool DUAL, bool RIGHT>
__device__ __forceinline__ void dpx_profile(uint32_t (&sv)[8], uint32_t tw, uint32_t qw, const DpxK& K)
{
    const uint32_t xr = tw ^ qw, ne = xr | (xr >> 1);             // bit 2c set <=> lane c mismatches
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t sh = k < 4 ? (ne << (7 - 2 * k)) : (ne << (15 - 2 * k));
        const uint32_t mm = prmt(sh, 0u, k < 4 ? 0xAA88u : 0xBB99u);   // 0xffff per mismatching half
        sv[k] = (mm & K.sMis) | (~mm & K.sMch);
    }
}
// the same for a task with wildcard bases: `amb` holds one bit per lane (target window in the high half,
// query window in the low half); a lane where either base is the wildcard scores sc_N (:130-134)
template <bool DUAL, bool RIGHT>
__device__ __forceinline__ void dpx_profile_amb(uint32_t (&sv)[8], uint32_t amb, const DpxK& K)
{
    const uint32_t bits = (amb | (amb >> 16)) & 0xffffu;
#pragma unroll
    for (int k = 0; k < 8; ++k) { const uint32_t lm = prmt(bits << (7 - k), 0u, 0x9988u); sv[k] = (K.sAmb & lm) | (sv[k] & ~lm); }
}

// ~x as x * -1 + -1: an IMAD (FMA pipe) instead of a LOP3 (ALU pipe).  m1 = 0xffffffff in a register the compiler cannot see through.
__device__ __forceinline__ uint32_t not_fma(uint32_t x, uint32_t m1)
{
#ifdef FSV_NOT_LOP3      // experiment: plain complement
    (void)m1; return ~x;
#else
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %2;" : "=r"(d) : "r"(x), "r"(m1));
    return d;
#endif
}

// The recurrence on the 16 lanes of one vector (:26-47, :171-196), words 7..0 so that word k-1 is
// still "old" when word k reads it.  XT0/VT0/X2T0 are the t-1 operands of word 0.
template <bool DUAL, bool TB, bool RIGHT>
__device__ __forceinline__ void dpx_cells(uint32_t (&U)[8], uint32_t (&V)[8], uint32_t (&X)[8], uint32_t (&Y)[8],
                                          uint32_t (&X2)[8], uint32_t (&Y2)[8], const uint32_t (&S)[8],
                                          uint32_t XT0, uint32_t VT0, uint32_t X2T0, const DpxK& K, uint32_t ALL1, uint4& tbo)
{
    using C = DpxConst<DUAL, RIGHT>;
    const uint32_t fE = both(C::cE), fF = both(C::cF), fE2 = both(C::cE2), fF2 = both(C::cF2);      // max(.,0) floors
    // min(x, code | bit) is `code` when x's value byte is 0 and `code | bit` otherwise: the continuation bit of
    // ksw2.h:116-118 lands at its final position (0x08 E, 0x10 F, 0x20 E~, 0x40 F~) with one VIMNMX each
    const uint32_t oE = both(0x08u | C::cE), oF = both(0x10u | C::cF), oE2 = both(0x20u | C::cE2), oF2 = both(0x40u | C::cF2);
    uint32_t tbw[8];
#pragma unroll
    for (int k = 7; k >= 0; --k) {
        const uint32_t xt1 = k ? X[k - 1] : XT0, vt1 = k ? V[k - 1] : VT0, x2t1 = DUAL ? (k ? X2[k - 1] : X2T0) : 0;
        const uint32_t ut = U[k];
        uint32_t nz, a, b, a2 = 0, b2 = 0;
        a = __vadd2(xt1, vt1);
        b = __vadd2(Y[k], ut);
        // z is kept COMPLEMENTED.  There is no 16x2 subtract on sm_100a (__vsub2 costs three instructions: LOP3 ~b, VIADD.16x2
        // +0x00010001, VIADD.16x2), but every value here is an int8 in the HIGH byte of its half with a known LOW byte, and ~ turns
        // min into max.  With nz = ~z carrying a low byte of 0xff (one LOP3: ~zk | 0x00ff00ff, which also drops the tie-break code):
        //   clamp:  ~min(z, c) = max(~z, ~c)                                  (signed and unsigned alike)
        //   u = z - v[t-1] = ~(nz + v[t-1]),  v = z - u = ~(nz + u)           (low bytes 0xff + 0 -> 0xff, no carry; ~ -> 0)
        //   q - z = (q | 1) + nz                                              (low bytes 1 + 0xff: the carry is the +1 of the two's complement)
        // and the two remaining complements are x * -1 + -1 on the FMA pipe (IMAD), which the ALU-heavy mix of this loop leaves idle:
        // 8 instructions per word for clamp, u, v, q - z, q2 - z instead of 10 (13 with __vsub2).
        uint32_t zk;
        if (DUAL) {
            a2 = __vadd2(x2t1, vt1);
            b2 = __vadd2(Y2[k], ut);
            zk = __vimax3_s16x2(__vimax3_s16x2(S[k], a, b), a2, b2);
            nz = __vmaxs2(~zk | 0x00ff00ffu, K.nClamp);
        } else {
            const uint32_t t1 = __vmaxs2(S[k], a);                 // signed (:179)
            zk = __vmaxs2(t1, b);                                  // d = b > z (signed compare, :180)
            nz = __vmaxu2(__vminu2(~t1 | 0x00ff00ffu, ~b | 0x00ff00ffu), K.nClamp);   // unsigned (:41-42)
        }
        U[k] = not_fma(__vadd2(nz, vt1), ALL1);
        V[k] = not_fma(__vadd2(nz, ut), ALL1);
        const uint32_t n1 = __vadd2(K.kQ1, nz);
        const uint32_t kq2 = K.kQ21;
        uint32_t fl = 0;
        if (RIGHT) {
            // right alignment sets a continuation bit when the gap value is >= 0 BEFORE the max with 0 (:212-218),
            // so the sum is kept and its sign bits are moved to 0x08 / 0x10 / 0x20 / 0x40
            const uint32_t pe = __vadd2(a, n1), pf = __vadd2(b, n1);
            const uint32_t xa = __vmaxs2(pe, fE), ya = __vmaxs2(pf, fF);
            fl = ((~pe >> 12) & 0x00080008u) | ((~pf >> 11) & 0x00100010u);
            if (DUAL) {
                const uint32_t n2 = __vadd2(kq2, nz);
                const uint32_t pe2 = __vadd2(a2, n2), pf2 = __vadd2(b2, n2);
                const uint32_t xa2 = __vmaxs2(pe2, fE2), ya2 = __vmaxs2(pf2, fF2);
                X[k] = __vadd2(xa, K.nQE); Y[k] = __vadd2(ya, K.nQE);
                X2[k] = __vadd2(xa2, K.nQE2); Y2[k] = __vadd2(ya2, K.nQE2);
                fl |= ((~pe2 >> 10) & 0x00200020u) | ((~pf2 >> 9) & 0x00400040u);
            } else { X[k] = xa; Y[k] = ya; }
        } else {
            const uint32_t xa = __viaddmax_s16x2(a, n1, fE), ya = __viaddmax_s16x2(b, n1, fF);
            if (DUAL) {
                const uint32_t n2 = __vadd2(kq2, nz);
                const uint32_t xa2 = __viaddmax_s16x2(a2, n2, fE2), ya2 = __viaddmax_s16x2(b2, n2, fF2);
                X[k] = __vadd2(xa, K.nQE); Y[k] = __vadd2(ya, K.nQE);
                X2[k] = __vadd2(xa2, K.nQE2); Y2[k] = __vadd2(ya2, K.nQE2);
                if (TB) fl = __vmins2(xa, oE) + __vmins2(ya, oF) + __vmins2(xa2, oE2) + __vmins2(ya2, oF2);
            } else {
                X[k] = xa; Y[k] = ya;
                if (TB) fl = __vmins2(xa, oE) + __vmins2(ya, oF);
            }
        }
        // flags | code in the low byte of each half, in ONE LOP3: (fl & m) | (zk & ~m).  The low 3 bits of fl (left: the sum of the codes,
        // < 8) are masked away, bits 3..7 of zk's low byte are 0 (codes < 8), and the high bytes (zk's value) are never looked at below
        // (inline PTX: the compiler splits the C expression into two LOP3s, one per immediate)
        if (TB) asm("lop3.b32 %0, %1, %2, 0x00780078, 0xE4;" : "=r"(tbw[k]) : "r"(fl), "r"(zk));      // 0xE4 = (a & c) | (b & ~c)
    }
    if (TB) {   // 16 traceback bytes in lane order (:195)
        const uint32_t a01 = prmt(tbw[0], tbw[1], 0x6240u), a23 = prmt(tbw[2], tbw[3], 0x6240u);
        const uint32_t a45 = prmt(tbw[4], tbw[5], 0x6240u), a67 = prmt(tbw[6], tbw[7], 0x6240u);
        tbo.x = prmt(a01, a23, 0x5410u); tbo.y = prmt(a45, a67, 0x5410u);
        tbo.z = prmt(a01, a23, 0x7632u); tbo.w = prmt(a45, a67, 0x7632u);
    }
}

// EXCL only tags a second copy of each kernel: it is launched with (almost) all of an SM's
// shared memory reserved, so that a CTA working on one of the few very long tasks has its SM to itself.
// TBM: 0 = score only, 1 = traceback, ties to the left (default), 2 = traceback, ties to the right (KSW_EZ_RIGHT),
// 3 = score only with KSW_EZ_APPROX_MAX (:270-286): no H[] at all, one cell is followed greedily.
// End of a segmented task (one thread): its traceback pages go back to the pool.
__device__ __forceinline__ void seg_release(const RunCtx& C, const DevTask& T, const int32_t* table)
{
    pool_free(C.pool, T.tb_pages, table);
}

// Spin until *flag == want (one thread).  Traps (the launch fails, nothing hangs) after twice the pool's stall limit: every
// wait of this kind ends when an earlier task's CIGAR is written, which takes a fraction of a second.
static __device__ __noinline__ void seg_spin(const int32_t* flag, int32_t want, int stall_ms, const char* what)
{
    unsigned ns = 128;
    const long long t0 = global_ns();
    while (__ldcg(flag) != want) {
        __nanosleep(ns);
        if (ns < 4096) ns <<= 1;
        else if (global_ns() - t0 > 2ll * stall_ms * 1000000ll) { printf("fsv segment stall: block %d waits for %s %d, sees %d\n", (int)blockIdx.x, what, want, __ldcg(flag)); __trap(); }
    }
    __threadfence();
}

// A segmented task takes ALL its traceback pages at once when its first segment starts (thread 0 of that CTA), in the order of
// the segment queue (tickets): a task either holds everything it needs or nothing, so the segments of the tasks in flight can
// always finish.  While it waits it holds a reserve that keeps NEW ordinary tasks from starting (running ones still grow).
static __device__ __noinline__ void seg_admit(const RunCtx& C, const DevTask& T, const SegTask& ST, int32_t* table, int32_t* ticket)
{
    StallWatch watch;
    seg_spin(ticket, ST.ticket, C.pool.stall_ms, "ticket");
    for (;;) {
        atomicMax(C.pool.reserve, T.tb_pages);
        if (pool_try_alloc(C.pool, T.tb_pages, table, true)) break;
        __nanosleep(4000);
        watch.poll(C.pool);
    }
    atomicExch(C.pool.reserve, 0);
    __threadfence();
    atomicExch(C.seg_admitted + T.seg_id, 1);
    atomicAdd(ticket, 1);
}

// SEG: the launch works on SEGMENTS of long tasks (fsv_common.cuh, DevSeg): P.Q holds segment indices.
template <bool DUAL, int TBM, int NW, bool EXCL = false, bool SEG = false>
__global__ void __launch_bounds__(NW * 32, DpxOcc<NW>::value) fsv_fill_dpx_kernel(const __grid_constant__ DpxParams P)
{
    constexpr int NT = NW * 32;
    constexpr bool TB = TBM == 1 || TBM == 2, RIGHT = TBM == 2, APPROX = TBM == 3;
    using KC = DpxConst<DUAL, RIGHT>;
    // one shared block, addressed from a single base register:
    //   edge slots [2 parities][NW] x 32 B : {x, v, x2, qw} of lane 15 of each warp's last vector, then its H and wildcard bit
    //   {tie key of antidiagonal j, CTA-wide max H of antidiagonal j+1} x 3 (rings over 3 antidiagonals: the pair is what thread 0
    //   resets with one store per iteration; INT32_MAX as a maximum = "stop": z-drop, or a cancelled segment); H[en0] / H[st0] rings; task index;
    //   traceback pages the task holds (thread 0's; fewer than tb_pages = a lazily growing task);
    //   APPROX: v[t*] and u[t*+1] of the followed cell, by antidiagonal parity
    constexpr uint32_t OFF_EDGE = 0, OFF_KM = 2 * NW * 32, OFF_HEN0 = OFF_KM + 24,
                       OFF_HST0 = OFF_HEN0 + 12, OFF_TASK = OFF_HST0 + 12, OFF_HELD = OFF_TASK + 4, OFF_APV = OFF_HELD + 4, OFF_APU = OFF_APV + 8, OFF_SEG = OFF_APU + 8, OFF_SCAN = OFF_SEG + 8,
                       SH_BYTES = OFF_SCAN + (SEG ? 4 * NW + 4 * NT : 0);      // SEG: scratch of the H re-anchoring scan
    __shared__ __align__(16) uint32_t sh_raw[(SH_BYTES + 15) / 16 * 4];
    uint32_t sb = (uint32_t)__cvta_generic_to_shared(sh_raw);
    const RunCtx& C = P.C;
    const DevScoring& sc = C.sc;
    int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifndef FSV_NO_OPAQUE
    // opaque to the compiler: otherwise it rebuilds these from S2R / the shared window base inside the fill loop
    // (10+ instructions per antidiagonal) instead of keeping four registers
    asm volatile("" : "+r"(tid), "+r"(lane), "+r"(warp), "+r"(sb));
#endif
    const unsigned FULL = 0xffffffffu;
    const DpxK& K = P.K;
    uint32_t ALL1 = 0xffffffffu;
    asm volatile("" : "+r"(ALL1));     // opaque: see not_fma
    const uint32_t extSel = DUAL ? 0xB391u : 0x4341u;                  // int8 (signed / unsigned) -> int16
    int32_t* table = C.page_tables + (int64_t)blockIdx.x * C.max_pages_per_task;
    int pending = -1, kept = 0;      // (thread 0) a task claimed but still waiting for memory; pages kept from the previous task
    int seg_redo = -1;          // SEG: a segment whose cold start did not reach its predecessor's state: this CTA runs it again FROM that state

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            int held = 0;
            if (SEG) sts32(sb + OFF_TASK, (uint32_t)(seg_redo >= 0 ? seg_redo : queue_take(P.Q, 0)));
            else sts32(sb + OFF_TASK, (uint32_t)next_task(C, P.Q, table, pending, true, held, kept));
            sts32(sb + OFF_KM + 8u * 2u + 4u, (uint32_t)INT32_MIN);      // the maximum slot of the first antidiagonal (the others are reset on the way)
            sts32(sb + OFF_HELD, (uint32_t)held);
        }
        __syncthreads();
        const int wi = (int)lds32(sb + OFF_TASK);      // task index, or (SEG) segment index
        if (wi < 0) return;
        // a segment of a long task: from its cold start, or (seg_redo) again from its predecessor's end state ("repair")
        DevSeg G{wi, 0, 1, 0, 0, 0};
        if (SEG) G = C.segs[wi];
        const int ti = G.task;
        const DevTask T = C.tasks[ti];
        const bool segmode = SEG;
        const bool repair = SEG && seg_redo >= 0;
        if (SEG) {
            const SegTask& STa = C.seg_tasks[T.seg_id];
            table = const_cast<int32_t*>(C.seg_tables) + STa.table_off;
            seg_redo = -1;
            if (!repair) {
                if (tid == 0) {
                    if (G.index == 0) seg_admit(C, T, STa, table, C.seg_ticket + P.seg_launch);
                    else seg_spin(C.seg_admitted + T.seg_id, 1, C.pool.stall_ms, "admission");
                }
                __syncthreads();
            }
        }
        const int rz = segmode ? (repair ? G.r_begin - 1 : G.r0) : 0;      // first antidiagonal computed (repair: the last one of the predecessor, whose state is loaded)
        const int r_own = segmode ? G.r_begin : 0;                         // first antidiagonal whose results count
        if (tid == 0 && C.timeline && (!SEG || G.index == 0)) C.timeline[2 * T.orig] = global_ns();
        const int qlen = T.qlen, tlen = T.tlen, w = T.w;
        const uint8_t* query = C.qarena + T.q_off;
        const uint8_t* target = C.tarena + T.t_off;
        // traceback rows live in pool pages: row r is in page r / rows_per_page
        int tb_rip = SEG ? r_own % T.rows_per_page : 0, tb_pg = SEG ? r_own / T.rows_per_page : 0;     // row inside the current page, page number (a segmented task's pages are static: segments may share one)
        uint8_t* tb_page = TB ? C.pool.base + (int64_t)(SEG ? __ldcg(table + tb_pg) : table[tb_pg]) * C.pool.page_bytes : nullptr;
        const int n_diag = qlen + tlen - 1;

        uint32_t U[8], V[8], X[8], Y[8], X2[8], Y2[8], S[8], Hr[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { U[k] = V[k] = X[k] = Y[k] = X2[k] = Y2[k] = S[k] = Hr[k] = 0; }
        int32_t Hb = 0;
        int Vt = tid - NT;          // forces the (re)arm path on the first antidiagonal
        uint32_t tw = 0, qw = 0;
        uint32_t amb = 0;           // wildcard bases of the two windows, one bit per lane: target << 16 | query (tasks with T.wild only)
        const bool wild_task = T.wild != 0;
        uint32_t qpre = 0; int qpre_r = -1;     // query base the lowest vector of the band needs at antidiagonal qpre_r
        EzState ez; ez.reset();     // complete only in warp 0 (the bookkeeping warp)
        int64_t cells = 0;
        int32_t hprev_keep = 0;     // H[en0-1] as last seen while that lane was inside the band
        // The ksw_extz_t bookkeeping of antidiagonal d is finished two iterations later (d+2), by warp 0
        // only: the maximum of d crosses the CTA through shared memory behind the ONE barrier of
        // iteration d, the tie-break key of the lanes holding it behind the barrier of iteration d+1.
        int stop_r = segmode ? G.r_end : n_diag;        // first antidiagonal that is not computed (band exhausted, :111; end of the segment)
        bool dropped = false;
        int32_t maxrun = 0;         // ez.max as every warp can track it (running maximum, ksw2.h:164)
        int32_t M1 = 0, M2 = 0;     // max H of antidiagonals r-1, r-2
        bool nt1 = false, nt2 = false;   // was the argmax of r-1 / r-2 needed?
        int32_t habs_p = INT32_MIN; bool act_p = false; int st0p = 0, en0p = 0;   // this thread at r-1
        int32_t H0 = 0; int ap_t = 0;   // APPROX: score and column of the one cell that is followed (last_H0_t, :271-283); every thread keeps them
        int s3 = 0;                 // r % 3

        if (SEG && repair) {
            // the predecessor's state after its last antidiagonal (snapshot slot 2*(index-1)): registers, sequence windows, and the
            // edge slots the first iteration reads; from here the run is the unsegmented one, in the predecessor's score frame
            const uint32_t* mine = C.seg_snap + C.seg_tasks[T.seg_id].snap_off + (int64_t)(2 * (G.index - 1)) * SEG_SNAP_WORDS + SEG_SNAP_HDR + tid;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                U[k] = __ldcg(mine + (0 + k) * NT); V[k] = __ldcg(mine + (8 + k) * NT); X[k] = __ldcg(mine + (16 + k) * NT); Y[k] = __ldcg(mine + (24 + k) * NT);
                X2[k] = __ldcg(mine + (32 + k) * NT); Y2[k] = __ldcg(mine + (40 + k) * NT); S[k] = __ldcg(mine + (48 + k) * NT); Hr[k] = __ldcg(mine + (56 + k) * NT);
            }
            Vt = (int)__ldcg(mine + 64 * NT); Hb = (int32_t)__ldcg(mine + 65 * NT); hprev_keep = (int32_t)__ldcg(mine + 66 * NT);
            const int rp = G.r_begin - 1, nb = Vt << 4;
            for (int c = 0; c < 16; ++c) {
                const int t = nb + c, j = rp - t;
                const uint32_t tbse = (t >= 0 && t < tlen) ? target[t] : 0, qbse = (j >= 0 && j < qlen) ? query[j] : 0;
                tw |= (tbse & 3u) << (2 * c); qw |= (qbse & 3u) << (2 * c);
                amb |= ((tbse > 3u ? 0x10000u : 0u) | (qbse > 3u ? 1u : 0u)) << c;
            }
            band_limits(rp, qlen, tlen, w, st0p, en0p);
            if (NW > 1 && lane == 31) {
                const uint32_t ea = sb + OFF_EDGE + (uint32_t)((rp & 1) * NW + warp) * 32u;
                sts128(ea, make_uint4(X[7], V[7], DUAL ? X2[7] : 0u, qw));
                sts32(ea + 16u, (uint32_t)(Hb + sext16(Hr[7] >> 16)));
                sts32(ea + 20u, amb);
            }
            __syncthreads();
        }
        // The antidiagonal loop exists twice, with and without the wildcard bookkeeping, chosen once per task:
        // left to itself the compiler predicates the `if (wild)` blocks, and predicated-off instructions still issue.
        auto fill_loop = [&](auto wild_c) {
        constexpr bool wild = decltype(wild_c)::value;
        for (int r = (SEG && repair) ? rz + 1 : rz;; ++r) {
            const int par = r & 1, ppar = par ^ 1;
            // a later segment of a task that z-dropped inside segment 0: stop through the same flag a z-drop uses (posted at
            // iteration r, seen by every thread in section (A) of iteration r+1)
            const int s3m1 = s3 == 0 ? 2 : s3 - 1, s3m2 = s3 == 2 ? 0 : s3 + 1;   // (r-1)%3, (r-2)%3
            // key of antidiagonal j: OFF_KM + 8j; maximum of antidiagonal j: OFF_KM + 8((j+2)%3) + 4
            const uint32_t a_key = sb + OFF_KM + 8u * (uint32_t)s3, a_key1 = sb + OFF_KM + 8u * (uint32_t)s3m1, a_key2 = sb + OFF_KM + 8u * (uint32_t)s3m2;
            const uint32_t a_max = a_key1 + 4u, a_max1 = a_key2 + 4u;
            if (SEG && segmode && G.index > 0 && (r & 255) == 0 && r > rz + 1 && tid == 0 && __ldcg(C.seg_cancel + T.seg_id) != 0)
                red_max_shared(a_max, INT32_MAX);
            // maximum of antidiagonal r-1 (every warp's red.max behind the last barrier), fetched early
            int32_t m_prev = INT32_MIN;
            if (!APPROX && r >= rz + 1) m_prev = (int32_t)lds32(a_max1);

            // ---- neighbour's lane 15 as it stood after the previous antidiagonal
            uint32_t nbX = __shfl_up_sync(FULL, X[7], 1), nbV = __shfl_up_sync(FULL, V[7], 1);
            uint32_t nbX2 = DUAL ? __shfl_up_sync(FULL, X2[7], 1) : 0;
            uint32_t nbQ = __shfl_up_sync(FULL, qw, 1);
            uint32_t nbA = 0;          // wildcard bit of the neighbour's lane 15 (query side)
            if (wild) {
                nbA = __shfl_up_sync(FULL, amb, 1);
                if (NW == 1) { const uint32_t a2 = __shfl_sync(FULL, amb, 31); if (lane == 0) nbA = a2; }
                else if (lane == 0 && r > rz) nbA = lds32(sb + OFF_EDGE + (uint32_t)(ppar * NW + (warp + NW - 1) % NW) * 32u + 20u);
                nbA = (nbA >> 15) & 1u;
            }
            if (NW == 1) {
                uint32_t a = __shfl_sync(FULL, X[7], 31), b = __shfl_sync(FULL, V[7], 31);
                uint32_t c2 = DUAL ? __shfl_sync(FULL, X2[7], 31) : 0, q2 = __shfl_sync(FULL, qw, 31);
                if (lane == 0) { nbX = a; nbV = b; nbX2 = c2; nbQ = q2; }
            } else if (lane == 0 && r > rz) {
                const int pw = (warp + NW - 1) % NW;
                const uint4 e = lds128(sb + OFF_EDGE + (uint32_t)(ppar * NW + pw) * 32u);
                nbX = e.x; nbV = e.y; nbX2 = e.z; nbQ = e.w;
            }

            if (APPROX) {
                // ---- approximate maximum (:270-286): antidiagonal d = r-1 posted v[t*] and u[t*+1] behind the last
                // barrier; every thread does the same scalar update, so all of them leave the loop together
                if (r >= 1) {
                    const int d = r - 1;
                    if (d >= stop_r) break;                                   // band exhausted (:111)
                    int st0d, en0d;
                    band_limits(d, qlen, tlen, w, st0d, en0d);
                    cells += en0d - st0d + 1;
                    const int32_t pv = (int32_t)lds32(sb + OFF_APV + 4u * (uint32_t)ppar), pu = (int32_t)lds32(sb + OFF_APU + 4u * (uint32_t)ppar);
                    if (d == 0) { H0 = pv + K.bias - K.r0_bias; ap_t = 0; }
                    else {
                        const bool in0 = ap_t >= st0d && ap_t <= en0d, in1 = ap_t + 1 >= st0d && ap_t + 1 <= en0d;
                        if (in0 && in1) { if (pv > pu) H0 += pv; else { H0 += pu; ++ap_t; } }
                        else if (in0) H0 += pv;
                        else { ++ap_t; H0 += pu; }
                    }
                    // extz2 tests the z-drop from the second antidiagonal on (:283 sits inside `if (r > 0)`), the dual variant on every one
                    if ((T.flag & FSV_EZ_APPROX_DROP) && (DUAL || d > 0) && ez.apply_zdrop(H0, d, ap_t, T.zdrop, sc.e_drop)) { dropped = true; break; }
                    if (d == n_diag - 1) { if (en0d == tlen - 1) ez.score = H0; break; }
                }
            }
            // ---- (A) antidiagonal d = r-2 is final: bookkeeping (ksw2_extz2_sse.c:262-269)
            if (!APPROX && r >= rz + 2) {
                if (m_prev == INT32_MAX) { dropped = true; break; }      // "stop" posted during the previous iteration: every thread agrees
                maxrun = max(maxrun, M2);
                const int d = r - 2;
                if (SEG && segmode) {
                    // a segment only RECORDS its own antidiagonals (scores relative to its own cold start): the
                    // bookkeeping is replayed over the whole task once every boundary has been checked
                    if (warp == 0 && d >= r_own && !dropped) {
                        int st0d, en0d;
                        band_limits(d, qlen, tlen, w, st0d, en0d);
                        int max_t = en0d;
                        const uint32_t bk = lds32(a_key2);
                        if (bk != 0) max_t = (int)((bk - 1u) & ((1u << 26) - 1u));
                        if (lane == 0) {
                            const int32_t hen0 = en0d == tlen - 1 ? (int32_t)lds32(sb + OFF_HEN0 + 4u * s3m2) : FSV_NEG_INF;
                            const int32_t hst0 = d - st0d == qlen - 1 ? (int32_t)lds32(sb + OFF_HST0 + 4u * s3m2) : FSV_NEG_INF;
                            C.seg_rec[C.seg_tasks[T.seg_id].rec_off + d] = make_int4(M2, max_t, hen0, hst0);
                        }
                        // segment 0 starts from the true state, so while the alignment is inside it the running maximum IS known:
                        // a z-drop here ends the task (the stitch's replay finds it again in the records) - the later segments
                        // would only be wasted work, and they are what a short batch would wait for: tell them to stop
                        if (G.index == 0 && ez.apply_zdrop(M2, d, max_t, T.zdrop, sc.e_drop)) {
                            dropped = true;
                            if (lane == 0) { red_max_shared(a_max, INT32_MAX); atomicExch(C.seg_cancel + T.seg_id, 1); }
                        }
                    }
                } else
                if (warp == 0 && !dropped) {
                    int st0d, en0d;
                    band_limits(d, qlen, tlen, w, st0d, en0d);
                    cells += en0d - st0d + 1;
                    int max_t = en0d;
                    if (nt2) {
                        const uint32_t bk = lds32(a_key2);
                        if (bk != 0) max_t = (int)((bk - 1u) & ((1u << 26) - 1u));
                    }
                    int32_t h_last = FSV_NEG_INF;
                    if (en0d == tlen - 1) {
                        h_last = (int32_t)lds32(sb + OFF_HEN0 + 4u * s3m2);
                        if (h_last > ez.mte) { ez.mte = h_last; ez.mte_q = d - round_en(en0d); }   // rounded en (:263-264)
                    }
                    if (d - st0d == qlen - 1) {
                        const int32_t h = (int32_t)lds32(sb + OFF_HST0 + 4u * s3m2);
                        if (h > ez.mqe) { ez.mqe = h; ez.mqe_t = st0d; }
                    }
                    if (ez.apply_zdrop(M2, d, max_t, T.zdrop, sc.e_drop)) { dropped = true; if (lane == 0) red_max_shared(a_max, INT32_MAX); }
                    else if (d == n_diag - 1 && en0d == tlen - 1) ez.score = h_last;            // H[tlen-1]
                }
                if (d == stop_r - 1) {     // every computed antidiagonal is final
                    if (NW > 1) __syncthreads(); else __syncwarp();
                    if ((int32_t)lds32(a_max) == INT32_MAX) dropped = true;      // posted during THIS iteration (an earlier one left the loop above)
                    break;
                }
            }
            // ---- (B) antidiagonal r-1: its maximum, and (only if observable, ksw2.h:164-174) who holds it
            if (!APPROX && r >= rz + 1 && r - 1 < stop_r) {
                const int32_t m = m_prev;
                M1 = m;
                const int32_t mr = max(maxrun, M2);      // ez.max once r-2 is accounted for
                nt1 = m > mr || (T.zdrop >= 0 && mr - m > T.zdrop) || (SEG && segmode);      // a segment cannot know: always
                if (nt1 && act_p && habs_p == m) {
                    // lanes of this vector that hold the maximum, as a 16-bit mask
                    const int base = Vt << 4;
                    const uint32_t mv = both((uint32_t)(m - Hb) & 0xffffu);
                    uint32_t ne = 0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) ne += __vminu2(Hr[k] ^ mv, 0x00010001u) << k;
                    uint32_t eq = ~((ne & 0xffu) | ((ne >> 8) & 0xff00u)) & lane_bits(st0p - base, en0p - base);
                    // order of the reference's scan (:231-260): lane en0, then four strided classes
                    // from st0 (first hit of each), then the scalar tail [en1, en0)
                    const int en1 = st0p + (en0p - st0p) / 4 * 4;
                    uint32_t key = 0xffffffffu;
                    if (en0p >= base && en0p <= base + 15 && ((eq >> (en0p - base)) & 1u)) key = 0;
                    else {
                        const int nb = min(max(en1 - base, 0), 16);          // lanes below en1
                        const uint32_t body = eq & ((1u << nb) - 1u), tail = eq & ~((1u << nb) - 1u);
                        const int sft = (base - st0p) & 3;
#pragma unroll
                        for (int i = 3; i >= 0; --i) {
                            const uint32_t mm = body & (0x1111u << ((i - sft) & 3));
                            if (mm) key = 1u + ((uint32_t)i << 26) + (uint32_t)(base + __ffs(mm) - 1);
                        }
                        if (!body && tail) key = 1u + (4u << 26) + (uint32_t)(base + __ffs(tail) - 1);
                    }
                    atom_min_shared(a_key1, key);
                }
            }
            if (!APPROX && tid == 0) sts64(a_key, 0xffffffffu, (uint32_t)INT32_MIN);      // key of r (posted at r+1), maximum of r+1

            // ---- (C) compute antidiagonal r
            int st0 = 0, en0 = -1;
            bool valid = false;
            // APPROX: the thread(s) owning columns t* and t*+1 post v[t*] and u[t*+1] of this antidiagonal as plain
            // differences (extz2: minus the q+e offset of its stored form), for the scalar update one iteration later
            auto approx_post = [&](int base) {
                const int l0 = ap_t - base, l1 = l0 + 1;
                if (l0 >= 0 && l0 < 16) { const uint32_t c = get_cell(V, l0); sts32(sb + OFF_APV + 4u * (uint32_t)par, (uint32_t)((DUAL ? (int)(int8_t)(c >> 8) : (int)((c >> 8) & 0xffu)) - K.bias)); }
                if (l1 >= 0 && l1 < 16) { const uint32_t c = get_cell(U, l1); sts32(sb + OFF_APU + 4u * (uint32_t)par, (uint32_t)((DUAL ? (int)(int8_t)(c >> 8) : (int)((c >> 8) & 0xffu)) - K.bias)); }
            };
            if (r < stop_r) {
                band_limits(r, qlen, tlen, w, st0, en0);
                if (st0 > en0) {
                    stop_r = r;
                    // a segment whose cold start already lies behind the end of the band (|qlen - tlen| > w) has nothing to compute;
                    // section (A) would never meet d == stop_r - 1 and run on past the task's records
                    if (SEG && segmode && r == rz) break;
                } else valid = true;
            }
            act_p = false;
            if (valid) {
                const int st = st0 & ~15, en = en0 | 15, st_ = st0 >> 4, en_ = en0 >> 4;      // round_st / round_en of 0 <= st0 <= en0 (:116)
                int32_t habs = INT32_MIN;      // this thread's best H over its in-band lanes
                // a warp whose 32 vectors all lie strictly inside the band takes the short path
                const bool inner = Vt > st_ && Vt < en_;
                if (__all_sync(FULL, inner)) {
                    const int base = Vt << 4;
                    qw = (qw << 2) | (nbQ >> 30);            // lane c now faces query[r - base - c]
                    dpx_profile<DUAL, RIGHT>(S, tw, qw, K);
                    if (wild) { amb = (amb & 0xffff0000u) | (((amb << 1) | nbA) & 0xffffu); dpx_profile_amb<DUAL, RIGHT>(S, amb, K); }
                    const uint32_t XT0 = prmt(nbX, X[7], 0x5432u), VT0 = prmt(nbV, V[7], 0x5432u);
                    const uint32_t X2T0 = DUAL ? prmt(nbX2, X2[7], 0x5432u) : 0;
                    uint4 o;
                    dpx_cells<DUAL, TB, RIGHT>(U, V, X, Y, X2, Y2, S, XT0, VT0, X2T0, K, ALL1, o);
                    if (TB && (!SEG || r >= r_own)) *reinterpret_cast<uint4*>(tb_page + (int64_t)tb_rip * T.pitch + (base - st)) = o;
                    if (APPROX) approx_post(base);
                    else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {            // H[t] += v[t] - qe (:239-241)
                        uint32_t dv = prmt(V[k], 0u, extSel);
                        if (!DUAL) dv = __vadd2(dv, K.nBias);
                        Hr[k] = __vadd2(Hr[k], dv);
                    }
                    const uint32_t pm = __vimax3_s16x2(__vimax3_s16x2(Hr[0], Hr[1], Hr[2]), __vimax3_s16x2(Hr[3], Hr[4], Hr[5]), __vmaxs2(Hr[6], Hr[7]));
                    const int mrel = max(sext16(pm), sext16(pm >> 16));
                    habs = Hb + mrel;
                    if (__builtin_expect((r & 31) == 31, 0)) {   // keep the relative scores small
                        Hb += mrel;
                        const uint32_t dd = both((uint32_t)mrel & 0xffffu);
#pragma unroll
                        for (int k = 0; k < 8; ++k) Hr[k] = __vsub2(Hr[k], dd);
                    }
                    act_p = true;
                    }
                } else {
                    // ---- general path: band edges, first row, profile overhang, idle and re-arming vectors
                    int32_t nbH = 0;
                    if (!APPROX) nbH = __shfl_up_sync(FULL, Hb + sext16(Hr[7] >> 16), 1);
                    if (APPROX) {}
                    else if (NW == 1) { int32_t h = __shfl_sync(FULL, Hb + sext16(Hr[7] >> 16), 31); if (lane == 0) nbH = h; }
                    else if (lane == 0 && r > rz) nbH = (int32_t)lds32(sb + OFF_EDGE + (uint32_t)(ppar * NW + (warp + NW - 1) % NW) * 32u + 16u);
                    bool rearmed = false;
                    if (__builtin_expect(Vt < st_, 0)) {            // a vector that fell below the band re-arms NT vectors to the right
                        Vt += NT;
                        while (Vt < st_) Vt += NT;
                        rearmed = true;
#pragma unroll
                        for (int k = 0; k < 8; ++k) { U[k] = K.gU; V[k] = K.gU; X[k] = K.gX; Y[k] = K.gY; X2[k] = K.gX2; Y2[k] = K.gY2; S[k] = K.sInit; Hr[k] = 0; }
                        if (SEG && segmode && rz > 0 && r == rz && Vt <= en_) {
                            // cold start of a segment: the reference's initial constants describe a surface that falls by q+e
                            // per lane, a cliff against the true one that the clamp lets erode only slowly; start FLAT
                            // instead (u = v = 0) so that the true diagonal takes over like a seed and the gap regimes
                            // spread from it at one lane per antidiagonal
                            const uint32_t flat = both(DUAL ? 0u : hi8(K.qe));
#pragma unroll
                            for (int k = 0; k < 8; ++k) { U[k] = flat; V[k] = flat; }
                        }
                        Hb = 0;
                        tw = 0; qw = 0; amb = 0;
                        const int nb = Vt << 4;
                        for (int c = 0; c < 16; ++c) {
                            const int t = nb + c, j = r - t;
                            uint32_t tbse = t < tlen ? target[t] : 0;
                            uint32_t qbse = (j >= 0 && j < qlen) ? query[j] : 0;
                            tw |= (tbse & 3u) << (2 * c);
                            qw |= (qbse & 3u) << (2 * c);
                            amb |= ((tbse > 3u ? 0x10000u : 0u) | (qbse > 3u ? 1u : 0u)) << c;
                        }
                    }
                    // Everything below is straight-line code (selects and lane masks instead of branches), so
                    // that the band-edge work of one or two threads is scheduled in between the cell updates
                    // of the whole warp instead of being serialised behind them; only what happens once per
                    // vector (re-arming above), once per task (first rows) or never (a negative carry) branches.
                    const int base = Vt << 4;
                    const bool active = Vt >= st_ && Vt <= en_;
                    const bool lo_edge = Vt == st_, hi_edge = Vt == en_;
                    {   // lane c now faces query[r - base - c]; the lowest vector of the band reads its new base
                        // from memory, one antidiagonal ahead (qpre) so that the load is off the critical path
                        // (qpre is the RAW byte: anything derived from it here would make the warp wait for the load at once)
                        if (lo_edge && !rearmed && qpre_r != r) { const int j = r - base; qpre = (j >= 0 && j < qlen) ? (uint32_t)query[j] : 0u; }
                        const uint32_t q0 = lo_edge ? (qpre & 3u) : (nbQ >> 30);
                        qw = rearmed ? qw : ((qw << 2) | q0);
                        if (wild && !rearmed) amb = (amb & 0xffff0000u) | (((amb << 1) | (lo_edge ? (qpre > 3u ? 1u : 0u) : nbA)) & 0xffffu);
                        int st0n, en0n;
                        band_limits(r + 1, qlen, tlen, w, st0n, en0n);
                        if (Vt == (st0n >> 4)) { const int j = r + 1 - base; qpre = (j >= 0 && j < qlen) ? (uint32_t)query[j] : 0u; qpre_r = r + 1; }
                    }
                    {   // profile stores (:126-140): whole 16-lane stores from st0, so the last one overhangs en0
                        const int store_end = st0 + ((en0 - st0) >> 4) * 16 + 15;
                        uint32_t sv[8];
                        dpx_profile<DUAL, RIGHT>(sv, tw, qw, K);
                        if (wild) dpx_profile_amb<DUAL, RIGHT>(sv, amb, K);
                        const uint32_t bits = lane_bits(min(st0 - base, 16), store_end - base);      // 0 below st_ and above the overhang
#pragma unroll
                        for (int k = 0; k < 8; ++k) { const uint32_t lm = word_mask(bits, k); S[k] = (sv[k] & lm) | (S[k] & ~lm); }
                    }
                    if (active) {
                        // operands of word 0: lane -1 is the neighbour's lane 15, lane 7 is our own word 7
                        uint32_t XT0 = prmt(nbX, X[7], 0x5432u), VT0 = prmt(nbV, V[7], 0x5432u);
                        uint32_t X2T0 = DUAL ? prmt(nbX2, X2[7], 0x5432u) : 0;
                        const int ce = en0 - base; // lane of en0 inside this vector (meaningful when hi_edge)
                        {   // carries of the lowest vector (:118-122)
                            // [last_st, last_en] = rounded range of the previous antidiagonal (:287), rebuilt from its exact limits
                            const bool inl = st > 0 && st - 1 >= (st0p & ~15) && st - 1 <= (en0p | 15);
                            uint32_t vfirst = 0;   // first column: v1 of an antidiagonal that starts at t = 0 (only while r <= w)
                            if (__builtin_expect(st == 0, 0)) {
                                if (DUAL) vfirst = hi8(r == 0 ? -K.qe : r < sc.long_thres ? -sc.e : r == sc.long_thres ? sc.long_diff : -sc.e2);
                                else vfirst = hi8(r ? sc.q : 0);
                            }
                            const uint32_t x1 = inl ? (XT0 & 0xffffu) : (K.gX & 0xffffu);
                            const uint32_t v1 = inl ? (VT0 & 0xffffu) : st > 0 ? (K.gU & 0xffffu) : vfirst;
                            const uint32_t x21 = inl ? (X2T0 & 0xffffu) : (K.gX2 & 0xffffu);
                            XT0 = lo_edge ? ((XT0 & 0xffff0000u) | x1) : XT0;
                            VT0 = lo_edge ? ((VT0 & 0xffff0000u) | v1) : VT0;
                            X2T0 = lo_edge ? ((X2T0 & 0xffff0000u) | x21) : X2T0;
                            if (!DUAL && lo_edge && ((x1 | v1) & 0x8000u)) {   // _mm_cvtsi32_si128(int8_t) sign-extends a negative carry into lanes 1..3 (:146-147)
                                if (x1 & 0x8000u) { X[0] = (X[0] & 0xffff0000u) | 0xff00u | KC::cE; X[1] = (X[1] & 0xffff0000u) | 0xff00u | KC::cE; X[2] = (X[2] & 0xffff0000u) | 0xff00u | KC::cE; }
                                if (v1 & 0x8000u) { V[0] = (V[0] & 0xffff0000u) | 0xff00u; V[1] = (V[1] & 0xffff0000u) | 0xff00u; V[2] = (V[2] & 0xffff0000u) | 0xff00u; }
                            }
                        }
                        if (__builtin_expect(en >= r, 0)) {             // first row (:123): only while r <= w
                            if ((r >> 4) == Vt) {
                                uint32_t eu;
                                if (DUAL) eu = hi8(r == 0 ? -K.qe : r < sc.long_thres ? -sc.e : r == sc.long_thres ? sc.long_diff : -sc.e2);
                                else eu = hi8(r ? sc.q : 0);
                                set_cell(U, r & 15, eu); set_cell(Y, r & 15, K.gY & 0xffffu);
                                if (DUAL) set_cell(Y2, r & 15, K.gY2 & 0xffffu);
                            }
                        }
                        // H[en0-1] as the reference's H[] holds it (:231): the value of the previous antidiagonal
                        // while that lane was inside the band, else the value it had when it left the band
                        if (!APPROX) {
                            const int32_t hleft = ce > 0 ? Hb + sext16(get_cell(Hr, (ce - 1) & 15)) : nbH;
                            hprev_keep = (hi_edge && r > 0 && en0 > 0 && en0 - 1 >= st0p) ? hleft : hprev_keep;
                        }

                        uint4 o;
                        dpx_cells<DUAL, TB, RIGHT>(U, V, X, Y, X2, Y2, S, XT0, VT0, X2T0, K, ALL1, o);
                        if (TB && (!SEG || r >= r_own)) *reinterpret_cast<uint4*>(tb_page + (int64_t)tb_rip * T.pitch + (base - st)) = o;
                        if (APPROX) approx_post(base);
                        else {
                        // exact max bookkeeping (:224-260): H[t] += v[t] - qe ; H[en0] from its left neighbour
                        const bool fix = hi_edge && (r == 0 || en0 > 0);
                        int32_t fixv;
                        {
                            const uint32_t u16 = r == 0 ? (V[0] & 0xffffu) : get_cell(U, ce & 15);        // r == 0: H[0] from v[0] (:262)
                            const int un = DUAL ? (int)(int8_t)(u16 >> 8) : (int)((u16 >> 8) & 0xffu);
                            fixv = (r == 0 ? -K.r0_bias : hprev_keep - K.bias) + un;                      // absolute
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            uint32_t dv = prmt(V[k], 0u, extSel);
                            if (!DUAL) dv = __vadd2(dv, K.nBias);
                            Hr[k] = __vadd2(Hr[k], dv);
                        }
                        {   // lane en0 takes the value derived from its left neighbour; a vector whose lane 0 is en0
                            // (or the very first cell) restarts its base there
                            Hb = (fix && (r == 0 || ce == 0)) ? fixv : Hb;
                            const uint32_t fv = both((uint32_t)(fixv - Hb) & 0xffffu);
                            const uint32_t fbit = fix ? (1u << (ce & 15)) : 0u;
#pragma unroll
                            for (int k = 0; k < 8; ++k) { const uint32_t lm = word_mask(fbit, k); Hr[k] = (fv & lm) | (Hr[k] & ~lm); }
                        }
                        const uint32_t inb = lane_bits(st0 - base, en0 - base);
                        uint32_t pm = 0x80008000u;
#pragma unroll
                        for (int k = 0; k < 8; ++k) { const uint32_t lm = word_mask(inb, k); pm = __vmaxs2(pm, (Hr[k] & lm) | (0x80008000u & ~lm)); }
                        if (hi_edge && en0 == tlen - 1) sts32(sb + OFF_HEN0 + 4u * s3, (uint32_t)(Hb + sext16(get_cell(Hr, en0 - base))));
                        if (lo_edge && r - st0 == qlen - 1) sts32(sb + OFF_HST0 + 4u * s3, (uint32_t)(Hb + sext16(get_cell(Hr, st0 - base))));
                        const int mrel = max(sext16(pm), sext16(pm >> 16));
                        habs = Hb + mrel;
                        if (__builtin_expect((r & 31) == 31, 0)) {   // keep the relative scores small
                            Hb += mrel;
                            const uint32_t dd = both((uint32_t)mrel & 0xffffu);
#pragma unroll
                            for (int k = 0; k < 8; ++k) Hr[k] = __vsub2(Hr[k], dd);
                        }
                        act_p = true;
                        }
                    }
                }
                habs_p = habs; st0p = st0; en0p = en0;
                if (SEG && segmode && G.index > 0 && r == r_own - 1) {
                    // End of the warm-up: the difference arrays have converged to the truth, the tracked H has not
                    // (every lane started from its own zero).  Rebuild it on this antidiagonal from ONE anchor, lane st0
                    // := 0, with the identity the reference's own H updates keep exactly for neighbouring in-band lanes:
                    // H[t] - H[t-1] = u[t] - v[t-1]  (:231,239-241).  From here on H is the truth minus one constant.
                    uint32_t vprev = __shfl_up_sync(FULL, V[7], 1);
                    if (NW > 1) {
                        if (lane == 31) sts32(sb + OFF_SCAN + 4u * (uint32_t)warp, V[7]);
                        __syncthreads();
                        if (lane == 0) vprev = lds32(sb + OFF_SCAN + 4u * (uint32_t)((warp + NW - 1) % NW));
                    } else { const uint32_t x = __shfl_sync(FULL, V[7], 31); if (lane == 0) vprev = x; }
                    const int base = Vt << 4;
                    int pre[16], run = 0;
#pragma unroll
                    for (int cc = 0; cc < 16; ++cc) {
                        const uint32_t v16 = cc == 0 ? (vprev >> 16) : get_cell(V, cc - 1), u16 = get_cell(U, cc);
                        const int vl = DUAL ? (int)(int8_t)(v16 >> 8) : (int)((v16 >> 8) & 0xffu), ul = DUAL ? (int)(int8_t)(u16 >> 8) : (int)((u16 >> 8) & 0xffu);
                        const int t = base + cc;
                        if (t > st0 && t <= en0) run += ul - vl;
                        pre[cc] = run;
                    }
                    const int idx = min(max(Vt - st_, 0), NT - 1);               // position of this vector in the band
                    sts32(sb + OFF_SCAN + 4u * NW + 4u * (uint32_t)idx, (uint32_t)run);
                    __syncthreads();
                    int off = 0;
                    for (int k = 0; k < idx; ++k) off += (int)lds32(sb + OFF_SCAN + 4u * NW + 4u * (uint32_t)k);
                    Hb = off;
#pragma unroll
                    for (int k = 0; k < 8; ++k) Hr[k] = ((uint32_t)pre[k] & 0xffffu) | ((uint32_t)pre[k + 8] << 16);
                    __syncthreads();
                }
                if (SEG && segmode && ((r == r_own - 1 && G.index > 0) || (r == G.r_end - 1 && G.index + 1 < G.count))) {
                    // state after this antidiagonal: the last own row (slot 2*index) or the end of the warm-up (2*(index-1)+1)
                    const int slot = r == G.r_end - 1 ? 2 * G.index : 2 * (G.index - 1) + 1;
                    uint32_t* sn = C.seg_snap + C.seg_tasks[T.seg_id].snap_off + (int64_t)slot * SEG_SNAP_WORDS;
                    uint32_t* mine = sn + SEG_SNAP_HDR + tid;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        mine[(0 + k) * NT] = U[k]; mine[(8 + k) * NT] = V[k]; mine[(16 + k) * NT] = X[k]; mine[(24 + k) * NT] = Y[k];
                        mine[(32 + k) * NT] = X2[k]; mine[(40 + k) * NT] = Y2[k]; mine[(48 + k) * NT] = S[k]; mine[(56 + k) * NT] = Hr[k];
                    }
                    mine[64 * NT] = (uint32_t)Vt; mine[65 * NT] = (uint32_t)Hb; mine[66 * NT] = (uint32_t)hprev_keep;
                    if (Vt == st_) { sn[0] = (uint32_t)(Hb + sext16(get_cell(Hr, st0 - (Vt << 4)))); sn[1] = 1u; }     // anchor: H of lane st0
                }
                // lane 15 of every warp's last vector, for the next antidiagonal
                if (NW > 1 && lane == 31) {
                    const uint32_t ea = sb + OFF_EDGE + (uint32_t)(par * NW + warp) * 32u;
                    sts128(ea, make_uint4(X[7], V[7], DUAL ? X2[7] : 0u, qw));
                    if (!APPROX) sts32(ea + 16u, (uint32_t)(Hb + sext16(Hr[7] >> 16)));
                    if (wild) sts32(ea + 20u, amb);
                }
                if (!APPROX) {
                    const int32_t wmax = __reduce_max_sync(FULL, habs);
                    if (lane == 0) red_max_shared(a_max, wmax);
                }
                if (TB && (!SEG || r >= r_own) && __builtin_expect(++tb_rip == T.rows_per_page, 0)) {
                    tb_rip = 0; ++tb_pg;
                    if (tb_pg < T.tb_pages) tb_page = C.pool.base + (int64_t)(SEG ? __ldcg(table + tb_pg) : table[tb_pg]) * C.pool.page_bytes;
                    // (thread 0 of a lazily growing task) one page ahead: the CTA reads table[tb_pg + 1] a whole page of antidiagonals from now
                    if (!SEG && tid == 0 && tb_pg + 1 < T.tb_pages) {
                        const int held = (int)lds32(sb + OFF_HELD);
                        if (held < tb_pg + 2) sts32(sb + OFF_HELD, (uint32_t)pool_lazy_grow(C.pool, C.slot_base + (int)blockIdx.x, table, held, tb_pg + 2));
                    }
                }
            }
            // ---- (D) the one barrier of the antidiagonal
            if (NW > 1) __syncthreads(); else __syncwarp();
            M2 = M1; nt2 = nt1;
            s3 = s3 == 2 ? 0 : s3 + 1;
        }
        };
        if (wild_task) fill_loop(std::true_type{}); else fill_loop(std::false_type{});
        if (!dropped && stop_r < n_diag) ez.zdropped = 1;      // band exhausted (:111-114)
        if (SEG && segmode) {
            // ---- end of a segment; the CTA that finishes the task's last one stitches the task together
            const SegTask ST = C.seg_tasks[T.seg_id];
            __threadfence();                 // this thread's traceback rows, records and snapshot words
            __syncthreads();
            if (tid == 0) {
                C.seg_foot[ST.first_seg + G.index] = stop_r < G.r_end ? stop_r : -1;      // the band ran out inside this segment
                if (repair) {
                    // this segment now continues its predecessor exactly: boundary index-1 holds by construction and the two share
                    // one score frame (header word 2 of the snapshot the cold start had left at slot 2*(index-1)+1)
                    uint32_t* hb = C.seg_snap + ST.snap_off + (int64_t)(2 * (G.index - 1) + 1) * SEG_SNAP_WORDS;
                    hb[2] = 1u;
                    __threadfence();
                    sts32(sb + OFF_SEG, 1u);
                } else {
                    __threadfence();
                    sts32(sb + OFF_SEG, atomicAdd(&C.seg_done[T.seg_id], 1) == G.count - 1 ? 1u : 0u);
                }
            }
            __syncthreads();
            if (!lds32(sb + OFF_SEG)) continue;
            __threadfence();
            // (1) warp 0 replays the ksw_extz_t bookkeeping (:262-269) over the records, 32 at a time, and finds the segment
            // the alignment ends in (z-drop, band exhausted, or the last one).  The records of segment k are the true ones if
            // the boundaries below k hold, which (2) checks afterwards: segments BEHIND the end (an extension that z-dropped
            // early ran on unrelated sequence there, where a cold start need not converge) are never looked at.
            EzState e2; e2.reset();
            int64_t cells2 = 0;
            if (warp == 0) {
                int32_t delta = 0;
                bool stop = false;
                int stop_seg = G.count - 1;
                for (int sgi = 0; sgi < G.count && !stop; ++sgi) {
                    const DevSeg Gs = C.segs[ST.first_seg + sgi];
                    const int foot = ((volatile int32_t*)C.seg_foot)[ST.first_seg + sgi];
                    const int end = foot >= 0 ? foot : Gs.r_end;
                    int4 rec_next = make_int4(0, 0, 0, 0);
                    if (Gs.r_begin + lane < end) rec_next = __ldcg(C.seg_rec + ST.rec_off + Gs.r_begin + lane);
                    for (int d0 = Gs.r_begin; d0 < end && !stop; d0 += 32) {
                        // one record per lane; ksw_apply_zdrop's running maximum (ksw2.h:160-176: a later record wins only if
                        // strictly greater) is an inclusive warp scan with the state so far as the earliest element, every lane
                        // tests its own record against the maximum BEFORE it, and the first lane that drops ends the replay.
                        // (the records of the next step are fetched while this one is scanned: the loop is one dependent load long otherwise)
                        const bool valid = d0 + lane < end;
                        const int4 rec = rec_next;
                        if (d0 + 32 + lane < end) rec_next = __ldcg(C.seg_rec + ST.rec_off + d0 + 32 + lane);
                        const int d = d0 + lane;
                        int st0d, en0d;
                        band_limits(valid ? d : 0, qlen, tlen, w, st0d, en0d);
                        const int32_t M = valid ? rec.x + delta : (int32_t)0x80000000;
                        const int mt = rec.y, mq = d - rec.y;
                        int32_t pm = M; int pt = mt, pq = mq;
#pragma unroll
                        for (int off = 1; off < 32; off <<= 1) {
                            const int32_t om = __shfl_up_sync(FULL, pm, off);
                            const int ot = __shfl_up_sync(FULL, pt, off), oq = __shfl_up_sync(FULL, pq, off);
                            if (lane >= off && !(pm > om)) { pm = om; pt = ot; pq = oq; }
                        }
                        int32_t bm = __shfl_up_sync(FULL, pm, 1); int bt = __shfl_up_sync(FULL, pt, 1), bq = __shfl_up_sync(FULL, pq, 1);
                        if (lane == 0 || !(bm > e2.max)) { bm = e2.max; bt = e2.max_t; bq = e2.max_q; }      // the maximum before this record
                        bool drops = false;
                        if (valid && !(M > bm) && mt >= bt && mq >= bq) {
                            const int tl = mt - bt, ql = mq - bq, l = tl > ql ? tl - ql : ql - tl;
                            drops = T.zdrop >= 0 && bm - M > T.zdrop + l * sc.e_drop;
                        }
                        const uint32_t dmask = __ballot_sync(FULL, drops);
                        const int first = dmask ? __ffs((int)dmask) - 1 : -1;
                        const int nproc = first >= 0 ? first + 1 : min(32, end - d0);      // records that count, the dropping one included
                        const bool inc = lane < nproc;
                        {   // maximum after the last counted record
                            const int32_t lm = __shfl_sync(FULL, pm, nproc - 1);
                            const int lt = __shfl_sync(FULL, pt, nproc - 1), lq = __shfl_sync(FULL, pq, nproc - 1);
                            if (lm > e2.max) { e2.max = (int32_t)((uint32_t)lm & 0x7fffffffu); e2.max_t = lt; e2.max_q = lq; }
                        }
                        cells2 += (int64_t)__reduce_add_sync(FULL, inc ? (uint32_t)(en0d - st0d + 1) : 0u);
                        {   // mte / mqe (:262-267): strictly greater, so the FIRST record holding the chunk's maximum
                            const int32_t he = (inc && en0d == tlen - 1 && rec.z != FSV_NEG_INF) ? rec.z + delta : FSV_NEG_INF;
                            const int32_t mx = __reduce_max_sync(FULL, he);
                            if (mx > e2.mte) {
                                const int L = __ffs((int)__ballot_sync(FULL, he == mx)) - 1;
                                e2.mte = mx; e2.mte_q = __shfl_sync(FULL, d - round_en(en0d), L);
                            }
                            const int32_t hs = (inc && d - st0d == qlen - 1 && rec.w != FSV_NEG_INF) ? rec.w + delta : FSV_NEG_INF;
                            const int32_t mx2 = __reduce_max_sync(FULL, hs);
                            if (mx2 > e2.mqe) {
                                const int L = __ffs((int)__ballot_sync(FULL, hs == mx2)) - 1;
                                e2.mqe = mx2; e2.mqe_t = __shfl_sync(FULL, st0d, L);
                            }
                            const uint32_t last = __ballot_sync(FULL, inc && lane != first && d == n_diag - 1 && en0d == tlen - 1);
                            if (last) e2.score = __shfl_sync(FULL, rec.z != FSV_NEG_INF ? rec.z + delta : FSV_NEG_INF, __ffs((int)last) - 1);
                        }
                        if (first >= 0) { e2.zdropped = 1; stop = true; }
                    }
                    if (!stop && foot >= 0) { e2.zdropped = 1; stop = true; }                     // band exhausted (:111-114)
                    if (stop) stop_seg = sgi;
                    if (!stop && sgi + 1 < G.count) {
                        const uint32_t* A = C.seg_snap + ST.snap_off + (int64_t)(2 * sgi) * SEG_SNAP_WORDS;
                        if (__ldcg(A + SEG_SNAP_WORDS + 2) == 0u) delta += (int32_t)__ldcg(A) - (int32_t)__ldcg(A + SEG_SNAP_WORDS);      // (a repaired segment keeps its predecessor's frame)
                    }
                }
                if (lane == 0) sts32(sb + OFF_SEG, (uint32_t)stop_seg);
            }
            __syncthreads();
            const int n_bound = (int)lds32(sb + OFF_SEG);      // boundaries 0 .. n_bound-1 carry the result
            __syncthreads();
            // (2) those boundaries, in order: the state segment b+1 reached from its cold start == the state segment b reached.
            // The first one that does not hold is repaired (its upper segment runs again from the true state, on this CTA),
            // then everything is looked at again: the records above it were in the wrong frame, so the replay may end elsewhere.
            int first_bad = -1;
            for (int b = 0; b < n_bound; ++b) {
                if (((volatile int32_t*)C.seg_foot)[ST.first_seg + b] >= 0) break;       // the alignment ends inside segment b
                const uint32_t* A = C.seg_snap + ST.snap_off + (int64_t)(2 * b) * SEG_SNAP_WORDS;
                const uint32_t* B = A + SEG_SNAP_WORDS;
                if (__ldcg(B + 2) != 0u) continue;                                         // repaired before: holds by construction
                bool bad = false;
                const int rb = C.segs[ST.first_seg + b].r_end - 1;
                int st0b, en0b;
                band_limits(rb, qlen, tlen, w, st0b, en0b);
                const int stv = round_st(st0b) >> 4, env = round_en(en0b) >> 4;
                if (__ldcg(A + 1) != 1u || __ldcg(B + 1) != 1u) bad = true;
                const int32_t ancA = (int32_t)__ldcg(A), ancB = (int32_t)__ldcg(B);
                const uint32_t* a = A + SEG_SNAP_HDR + tid; const uint32_t* bq = B + SEG_SNAP_HDR + tid;
                const int va = (int)__ldcg(a + 64 * NT), vb = (int)__ldcg(bq + 64 * NT);
                if (va != vb) bad = true;
                if (va >= stv) for (int k = 48; k < 56; ++k) if (__ldcg(a + k * NT) != __ldcg(bq + k * NT)) bad = true;     // s, also of vectors above the band
                if (va >= stv && va <= env) {
                    for (int k = 0; k < 48; ++k) if (__ldcg(a + k * NT) != __ldcg(bq + k * NT)) bad = true;             // u v x y x2 y2
                    const int32_t hba = (int32_t)__ldcg(a + 65 * NT), hbb = (int32_t)__ldcg(bq + 65 * NT);
                    for (int c = 0; c < 16; ++c) {                                                                   // H relative to the anchor lane
                        const int t = (va << 4) + c;
                        if (t < st0b || t > en0b) continue;
                        const uint32_t wa = __ldcg(a + (56 + (c & 7)) * NT), wb = __ldcg(bq + (56 + (c & 7)) * NT);
                        const int32_t ha = hba + sext16(c & 8 ? wa >> 16 : wa) - ancA, hb = hbb + sext16(c & 8 ? wb >> 16 : wb) - ancB;
                        if (ha != hb) bad = true;
                    }
                }
                if (__syncthreads_or(bad)) { first_bad = b; break; }
            }
            if (first_bad >= 0) {
                if (tid == 0) atomicAdd(&C.seg_done[T.seg_id], 1 << 20);      // counted by the host (fsv_stats.segment_fallbacks = repaired segments)
                seg_redo = ST.first_seg + first_bad + 1;
                continue;
            }
            if (warp == 0) finish_task(C, T, table, e2, cells2, TB);
            __syncthreads();
            if (tid == 0) {
                seg_release(C, T, table);
                if (C.timeline) C.timeline[2 * T.orig + 1] = global_ns();
            }
            continue;
        }

        __syncthreads();             // every traceback row is written
        if (warp == 0) finish_task(C, T, table, ez, cells, TB);
        __syncthreads();
        if (tid == 0) {
            const bool lazy = C.pool.lazy && C.pool.lazy_min_pages > 0 && T.tb_pages >= C.pool.lazy_min_pages;      // as task_pages decided
            if (!SEG) {
                const int held = (int)lds32(sb + OFF_HELD);
                if (!lazy && pool_may_keep(C.pool, held)) kept = held;      // table[0 .. held) stays with this CTA for its next task
                else pool_free(C.pool, held, table, lazy ? C.slot_base + (int)blockIdx.x : -1);
            }
            if (C.timeline) C.timeline[2 * T.orig + 1] = global_ns();
        }
    }
}

// ---------------------------------------------------------------------------
// host side

// warps a task needs: one thread per band vector plus the profile-overhang vector
inline int dpx_warps_needed(const DevTask& t) { return (t.pitch / 16 + 1 + 31) / 32; }

inline int dpx_class_of(int warps)
{
    if (warps <= 1) return 1;
    if (warps <= 2) return 2;
    if (warps <= 4) return 4;
    if (warps <= 6) return 6;
    if (warps <= 8) return 8;
    return 0;
}

// `has_wild` = some base of the task is not A/C/G/T (code > 3)
inline bool dpx_supports(const DevScoring& sc, const DevTask& t, bool has_wild)
{
    (void)has_wild;            // wildcard bases: one extra bit per lane in the kernel (T.wild)
    if (t.kind != 1) return false;
    if (t.flag & FSV_EZ_GENERIC_SC) return false;
    if ((t.flag & FSV_EZ_APPROX_MAX) && !(t.flag & FSV_EZ_SCORE_ONLY)) return false;      // approximate maximum: score-only variant (TBM 3)
    if (sc.m != 5) return false;
    return dpx_class_of(dpx_warps_needed(t)) != 0;
}

constexpr int DPX_EXCL_SMEM = 226 * 1024;   // dynamic shared memory reserved by an exclusive CTA: nothing else fits on its SM

// CTAs to launch for n_tasks tasks of the NW-warp class (persistent CTAs)
template <bool DUAL, int TBM, int NW>
inline int dpx_grid_one(int sm_count, int n_tasks)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fsv_fill_dpx_kernel<DUAL, TBM, NW, false>, NW * 32, 0) != cudaSuccess) { cudaGetLastError(); per_sm = 1; }
    if (per_sm < 1) per_sm = 1;
    return std::max(1, std::min(n_tasks, sm_count * per_sm));
}

template <bool DUAL, int TBM, int NW>
inline int dpx_launch_one(cudaStream_t stream, int grid, bool excl, const DpxParams& P, std::string* err)
{
    cudaError_t e = cudaSuccess;
    DpxParams PK = P;
    PK.K = DpxConst<DUAL, TBM == 2>(P.C.sc);
    if (excl) {
        auto kern = fsv_fill_dpx_kernel<DUAL, TBM, NW, true>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DPX_EXCL_SMEM);
        if (e == cudaSuccess) kern<<<grid, NW * 32, DPX_EXCL_SMEM, stream>>>(PK);
    } else {
        fsv_fill_dpx_kernel<DUAL, TBM, NW, false><<<grid, NW * 32, 0, stream>>>(PK);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { if (err) *err = cudaGetErrorString(e); cudaGetLastError(); return FSV_ERR_CUDA; }
    return FSV_OK;
}

// segments of long tasks (TBM 1 only): persistent CTAs over the segment queue
template <bool DUAL, int NW>
inline int dpx_launch_seg_one(cudaStream_t stream, int sm_count, int n_segs, const DpxParams& P, std::string* err)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fsv_fill_dpx_kernel<DUAL, 1, NW, false, true>, NW * 32, 0) != cudaSuccess) { cudaGetLastError(); per_sm = 1; }
    const int grid = std::max(1, std::min(n_segs, sm_count * std::max(per_sm, 1)));
    DpxParams PK = P;
    PK.K = DpxConst<DUAL, false>(P.C.sc);
    fsv_fill_dpx_kernel<DUAL, 1, NW, false, true><<<grid, NW * 32, 0, stream>>>(PK);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { if (err) *err = cudaGetErrorString(e); return FSV_ERR_CUDA; }
    return FSV_OK;
}
template <bool DUAL>
inline int dpx_launch_seg_nw(cudaStream_t stream, int sm_count, int nw, int n_segs, const DpxParams& P, std::string* err)
{
    switch (nw) {
        case 1: return dpx_launch_seg_one<DUAL, 1>(stream, sm_count, n_segs, P, err);
        case 2: return dpx_launch_seg_one<DUAL, 2>(stream, sm_count, n_segs, P, err);
        case 4: return dpx_launch_seg_one<DUAL, 4>(stream, sm_count, n_segs, P, err);
        case 6: return dpx_launch_seg_one<DUAL, 6>(stream, sm_count, n_segs, P, err);
        case 8: return dpx_launch_seg_one<DUAL, 8>(stream, sm_count, n_segs, P, err);
    }
    return FSV_ERR_INVALID;
}

#define FSV_DPX_DISPATCH(CALL)                                                    \
    switch (nw) {                                                                 \
        case 1: return CALL(1);                                                   \
        case 2: return CALL(2);                                                   \
        case 4: return CALL(4);                                                   \
        case 6: return CALL(6);                                                   \
        case 8: return CALL(8);                                                   \
    }

template <bool DUAL, int TBM>
inline int dpx_grid_nw(int sm_count, int nw, int n_tasks)
{
#define FSV_G(N) dpx_grid_one<DUAL, TBM, N>(sm_count, n_tasks)
    FSV_DPX_DISPATCH(FSV_G)
#undef FSV_G
    return 1;
}
template <bool DUAL, int TBM>
inline int dpx_launch_nw(cudaStream_t stream, int nw, int grid, bool excl, const DpxParams& P, std::string* err)
{
#define FSV_L(N) dpx_launch_one<DUAL, TBM, N>(stream, grid, excl, P, err)
    FSV_DPX_DISPATCH(FSV_L)
#undef FSV_L
    return FSV_ERR_INVALID;
}

// One launch per (warps-per-task class, traceback mode 0/1/2, exclusive or not).  The six (DUAL, TBM) families
// are compiled in separate translation units (fsv_dpx_variant.cu, built in parallel); this is what they export.
#define FSV_DPX_FAMILY(D, T)                                                                                           \
    int dpx_grid_##D##T(int sm_count, int nw, int n_tasks);                                                           \
    int dpx_launch_##D##T(cudaStream_t stream, int nw, int grid, bool excl, const DpxParams& P, std::string* err);
int dpx_launch_seg_0(cudaStream_t stream, int sm_count, int nw, int n_segs, const DpxParams& P, std::string* err);
int dpx_launch_seg_1(cudaStream_t stream, int sm_count, int nw, int n_segs, const DpxParams& P, std::string* err);
FSV_DPX_FAMILY(0, 0) FSV_DPX_FAMILY(0, 1) FSV_DPX_FAMILY(0, 2) FSV_DPX_FAMILY(0, 3) FSV_DPX_FAMILY(1, 0) FSV_DPX_FAMILY(1, 1) FSV_DPX_FAMILY(1, 2) FSV_DPX_FAMILY(1, 3)
#undef FSV_DPX_FAMILY

inline int dpx_grid(int sm_count, bool dual, int tbm, int nw, int n_tasks)
{
    if (dual) return tbm == 3 ? dpx_grid_13(sm_count, nw, n_tasks) : tbm == 2 ? dpx_grid_12(sm_count, nw, n_tasks) : tbm ? dpx_grid_11(sm_count, nw, n_tasks) : dpx_grid_10(sm_count, nw, n_tasks);
    return tbm == 3 ? dpx_grid_03(sm_count, nw, n_tasks) : tbm == 2 ? dpx_grid_02(sm_count, nw, n_tasks) : tbm ? dpx_grid_01(sm_count, nw, n_tasks) : dpx_grid_00(sm_count, nw, n_tasks);
}
inline int dpx_launch(cudaStream_t stream, bool dual, int tbm, int nw, int grid, bool excl, const DpxParams& P, std::string* err)
{
    if (dual) return tbm == 3 ? dpx_launch_13(stream, nw, grid, excl, P, err) : tbm == 2 ? dpx_launch_12(stream, nw, grid, excl, P, err) : tbm ? dpx_launch_11(stream, nw, grid, excl, P, err) : dpx_launch_10(stream, nw, grid, excl, P, err);
    return tbm == 3 ? dpx_launch_03(stream, nw, grid, excl, P, err) : tbm == 2 ? dpx_launch_02(stream, nw, grid, excl, P, err) : tbm ? dpx_launch_01(stream, nw, grid, excl, P, err) : dpx_launch_00(stream, nw, grid, excl, P, err);
}

}  // namespace fsv
