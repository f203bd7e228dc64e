// fsv_peaks.cuh — dependency-free integer-issue microbenchmarks (roofline denominators).
//
// The fill kernels are integer-pipe bound, so the roofline peak is the
// MEASURED issue rate of the instructions they are made of, in lane-ops per second
// (one lane-op = one 32-bit lane executing one instruction; a 16x2 DPX instruction
// updates two DP cells per lane-op).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fsv {

// kind 0: VIADD.16x2            (packed add)
// kind 1: VIMNMX3.S16x2         (packed 3-input max)
// kind 2: VIADDMNMX.S16x2       (packed add+max)
// kind 3: LOP3                  (logic)
// kind 4: IMAD (fma pipe)       (32-bit multiply-add)
// kind 5: mix VIADD/VIMNMX3/VIADDMNMX/LOP3/IMAD in the proportions of the DPX fill kernel
// kind 6: PRMT                  (byte permute)
// kind 7: alternating VIMNMX3.S16x2 + IMAD (both pipes)
template <int KIND>
__global__ void __launch_bounds__(256) fsv_peak_kernel(uint32_t* out, int iters, uint32_t seed)
{
    constexpr int N = 8;   // independent chains per thread
    uint32_t a[N], b = seed | 0x00010001u, c = seed * 2654435761u;
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] = seed + threadIdx.x * 977u + i * 131u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (KIND == 0) a[i] = __vadd2(a[i], b);
                else if (KIND == 1) a[i] = __vimax3_s16x2(a[i], b, c);
                else if (KIND == 2) a[i] = __viaddmax_s16x2(a[i], b, c);
                else if (KIND == 3) a[i] = (a[i] & b) ^ c;
                else if (KIND == 4) a[i] = a[i] * b + c;
                else if (KIND == 6) a[i] = __byte_perm(a[i], b, 0x2541);
                else if (KIND == 7) { if (i & 1) a[i] = a[i] * b + c; else a[i] = __vimax3_s16x2(a[i], b, c); }
                else {
                    switch ((k + i) % 5) {
                        case 0: a[i] = __vadd2(a[i], b); break;
                        case 1: a[i] = __vimax3_s16x2(a[i], b, c); break;
                        case 2: a[i] = __viaddmax_s16x2(a[i], b, c); break;
                        case 3: a[i] = (a[i] & b) ^ c; break;
                        default: a[i] = __vsub2(a[i], c); break;
                    }
                }
            }
            b += 0x00010001u;
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) s ^= a[i];
    if (s == 0x12345678u) out[threadIdx.x] = s;   // keep the chains alive
}

constexpr int PEAK_OPS_PER_ITER = 8 * 8;   // per thread per iteration (the `b +=` is not counted)

}  // namespace fsv
