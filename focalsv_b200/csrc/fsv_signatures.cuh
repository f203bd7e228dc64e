// fsv_signatures.cuh — DEL / INS signatures straight from the device-resident CIGARs of a batch.
//
// The step that follows the alignment in the reference: extract_sig_from_cigar
// (focalsv/4_sv_calling/Dippav/extract_contig_signature_CCS.py:14-127) walks each aligned contig's CIGAR,
// keeps D / I operations of at least min_svlen, and folds neighbouring signatures of the same contig left to
// right (INS: both > 250 within 250 bp, both > 320 within 380 bp, both > 100 within 250 bp, :72-88; DEL: both
// > 150 starting within 150 bp, :103-111).  Doing it here means only the signatures (32 B each) cross PCIe
// instead of every CIGAR word.  One thread per task; the fold only ever changes the LAST signature of each
// list, so the walk streams: a running DEL and a running INS are emitted when the next one does not merge.
// Output order per task: its DELs, then its INSs (the order hook.signatures returns).
#pragma once
#include "fsv_common.cuh"

namespace fsv {

struct SigRun { long long pos; int svlen, read_start, read_end; bool live; };

// WRITE = false: count only.  Returns (n_del, n_ins) through the references.
template <bool WRITE>
__device__ inline void sig_walk(const uint32_t* cig, int n_cigar, long long ref_start, int min_svlen, int task,
                                fsv_signature* out_del, fsv_signature* out_ins, int& n_del, int& n_ins)
{
    long long ro = ref_start;
    int co = 0, nd = 0, ni = 0;
    const int hard = (n_cigar > 0 && (cig[0] & 0xfu) == 5u) ? (int)(cig[0] >> 4) : 0;      // :24-26
    SigRun D{0, 0, 0, 0, false}, I{0, 0, 0, 0, false};
    auto emit = [&](const SigRun& s, int svtype, fsv_signature* out, int& n) {
        if (WRITE) { fsv_signature r; r.task = task; r.svtype = svtype; r.pos = s.pos; r.svlen = s.svlen; r.read_start = s.read_start; r.read_end = s.read_end; r.pad_ = 0; out[n] = r; }
        ++n;
    };
    for (int k = 0; k < n_cigar; ++k) {
        const uint32_t w = cig[k];
        const int op = (int)(w & 0xfu), len = (int)(w >> 4);
        if (op == 0) { ro += len; co += len; }                         // :34-36
        else if (op == 4) co += len;                                     // :37-38
        else if (op == 2) {                                              // :39-42
            if (len >= min_svlen) {
                const SigRun s{ro, len, co + hard, co + hard + 1, true};
                if (D.live && D.svlen > 150 && s.svlen > 150 && llabs(s.pos - D.pos) < 150) {      // :103-111, merge_two_del :64-70
                    D.svlen = (int)(s.pos + s.svlen - D.pos); D.read_end = D.read_start + 1;
                } else { if (D.live) emit(D, 0, out_del, nd); D = s; }
            }
            ro += len;
        } else if (op == 1) {                                            // :43-46
            if (len >= min_svlen) {
                const SigRun s{ro, len, co + hard, co + hard + len, true};
                const long long near = llabs(s.pos - I.pos);
                if (I.live && ((I.svlen > 250 && s.svlen > 250 && near < 250) || (I.svlen > 320 && s.svlen > 320 && near < 380) ||
                               (I.svlen > 100 && s.svlen > 100 && near < 250))) {                  // :72-88, merge_two_ins :56-62
                    I.read_end = s.read_end; I.svlen = I.read_end - I.read_start;
                } else { if (I.live) emit(I, 1, out_ins, ni); I = s; }
            }
            co += len;
        }
    }
    if (D.live) emit(D, 0, out_del, nd);
    if (I.live) emit(I, 1, out_ins, ni);
    n_del = nd; n_ins = ni;
}

// counts[2*i] = DEL signatures of task i, counts[2*i+1] = INS
__global__ void fsv_sig_count_kernel(const fsv_result* results, const uint32_t* cigar, const long long* ref_start, int n,
                                     int min_svlen, int32_t* counts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fsv_result R = results[i];
    int nd, ni;
    sig_walk<false>(cigar + R.cigar_off, R.n_cigar, ref_start ? ref_start[i] : 0, min_svlen, i, nullptr, nullptr, nd, ni);
    counts[2 * i] = nd; counts[2 * i + 1] = ni;
}

// offsets[i] = index of task i's first signature in `out` (host prefix sum of the counts)
__global__ void fsv_sig_write_kernel(const fsv_result* results, const uint32_t* cigar, const long long* ref_start, int n,
                                     int min_svlen, const int32_t* counts, const long long* offsets, fsv_signature* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fsv_result R = results[i];
    int nd, ni;
    fsv_signature* o = out + offsets[i];
    sig_walk<true>(cigar + R.cigar_off, R.n_cigar, ref_start ? ref_start[i] : 0, min_svlen, i, o, o + counts[2 * i], nd, ni);
}

}  // namespace fsv
