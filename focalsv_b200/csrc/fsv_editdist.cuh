// fsv_editdist.cuh — batched global (NW) unit-cost edit distance: the second DP of the pipeline.
//
// What the reference computes with edlib.align(seq1, seq2)["editDistance"] when it de-duplicates INS alleles
// (focalsv/4_sv_calling/Dippav/remove_redundancy.py:57-63, remove_redundancy_region_based.py:123-128; default edlib
// mode NW, k = -1, i.e. the plain Levenshtein distance).  The value is unique, so parity is exact by definition.
//
// Algorithm: Myers' bit-vector recurrence in Hyyro's block formulation (vertical deltas Pv/Mv of 64 pattern rows per
// block, horizontal carry hin/hout in {-1,0,+1} between blocks).  B200 shape: ONE WARP PER PAIR; lane b owns
// block b of a strip of 32 blocks (2048 pattern rows) with its Pv, Mv and the match masks of at most 8 symbols in
// registers; the blocks of a strip run as a skewed wavefront (lane b is at text column s - b at step s), so the
// carry and the text symbol travel down the lanes with ONE shuffle per step.  Patterns longer than a strip take
// several passes; the carry out of a strip's last block is kept per text column (1 byte) in global memory and is
// the carry into lane 0 of the next strip.  The distance is m + sum of the carries out of the last row.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/focalsv_cuda.h"

namespace fsv {

constexpr int ED_MAX_SYMBOLS = 8;

struct EdParams {
    const uint8_t* a;            // patterns (symbol codes 0..K-1 after the host's remap)
    const uint8_t* b;            // texts
    const fsv_pair* pairs;
    const int32_t* order;        // pairs sorted by work, largest first
    int32_t n;
    int32_t* dist;               // per pair (caller order)
    int8_t* carry;               // per warp: max_text_len bytes
    int64_t carry_pitch;
    unsigned int* cursor;        // next pair to take
};

__global__ void __launch_bounds__(128) fsv_edit_distance_kernel(const EdParams P)
{
    const int lane = threadIdx.x & 31;
    const int warp_global = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    int8_t* carry = P.carry + (int64_t)warp_global * P.carry_pitch;
    const unsigned FULL = 0xffffffffu;
    for (;;) {
        unsigned int k = 0;
        if (lane == 0) k = atomicAdd(P.cursor, 1u);
        k = __shfl_sync(FULL, k, 0);
        if (k >= (unsigned)P.n) return;
        const int pi = P.order[k];
        const fsv_pair pr = P.pairs[pi];
        const int m = pr.a_len, n = pr.b_len;
        if (m <= 0 || n <= 0) { if (lane == 0) P.dist[pi] = m > 0 ? m : (n > 0 ? n : 0); continue; }
        const uint8_t* pat = P.a + pr.a_off;
        const uint8_t* txt = P.b + pr.b_off;
        const int n_blocks = (m + 63) >> 6;
        int total = m;                                   // D[m][0]; lane of the last block accumulates the last row
        for (int strip0 = 0; strip0 < n_blocks; strip0 += 32) {
            const int blk = strip0 + lane;               // this lane's block
            const bool have = blk < n_blocks;
            const bool last_blk = blk == n_blocks - 1;
            const int rows = have ? min(64, m - (blk << 6)) : 0;
            const int nb_strip = min(32, n_blocks - strip0);
            // match masks of this block
            unsigned long long peq[ED_MAX_SYMBOLS];
#pragma unroll
            for (int s = 0; s < ED_MAX_SYMBOLS; ++s) peq[s] = 0ull;
            for (int i = 0; i < rows; ++i) {
                const int c = pat[(blk << 6) + i] & (ED_MAX_SYMBOLS - 1);
#pragma unroll
                for (int s = 0; s < ED_MAX_SYMBOLS; ++s) if (s == c) peq[s] |= 1ull << i;
            }
            unsigned long long Pv = ~0ull, Mv = 0ull;
            const int out_bit = (rows > 0 ? rows : 64) - 1;   // row whose horizontal delta leaves this block
            int msg = 0;                                 // (symbol << 2) | (hout + 1) handed to the next lane
            const int steps = n + nb_strip - 1;
            for (int s = 0; s < steps; ++s) {
                int in = __shfl_up_sync(FULL, msg, 1);
                const int j = s - lane;                  // this lane's text column at this step
                if (lane == 0) {
                    const int jj = j < n ? j : n - 1;
                    const int hin0 = strip0 == 0 ? 1 : (int)carry[jj];
                    in = ((int)(txt[jj] & (ED_MAX_SYMBOLS - 1)) << 2) | (hin0 + 1);
                }
                if (have && j >= 0 && j < n) {
                    const int hin = (in & 3) - 1, c = in >> 2;
                    unsigned long long Eq = 0ull;
#pragma unroll
                    for (int q = 0; q < ED_MAX_SYMBOLS; ++q) Eq = q == c ? peq[q] : Eq;
                    const unsigned long long Xv = Eq | Mv;
                    if (hin < 0) Eq |= 1ull;
                    const unsigned long long Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
                    unsigned long long Ph = Mv | ~(Xh | Pv);
                    unsigned long long Mh = Pv & Xh;
                    const int hout = (int)((Ph >> out_bit) & 1ull) - (int)((Mh >> out_bit) & 1ull);
                    Ph <<= 1; Mh <<= 1;
                    if (hin < 0) Mh |= 1ull; else if (hin > 0) Ph |= 1ull;
                    Pv = Mh | ~(Xv | Ph);
                    Mv = Ph & Xv;
                    msg = (c << 2) | (hout + 1);
                    if (last_blk) total += hout;
                    else if (lane == 31) carry[j] = (int8_t)hout;      // carry into the next strip
                }
            }
            __syncwarp();
        }
        // the lane that owned the last block holds the distance
        const int owner = (n_blocks - 1) & 31;
        total = __shfl_sync(FULL, total, owner);
        if (lane == 0) P.dist[pi] = total;
    }
}

}  // namespace fsv
