// fsv_chain.cu — host side of row f2: seeding, chaining and decomposition of one (contig, window) pair into
// the small DP tasks minimap2 would hand to ksw2, so that a 200 kb x 200 kb pair costs a few hundred fills of a few
// hundred bases plus one rectangular fill per structural variant instead of one band-3001 DP over 1.2e9 cells.
//
// What it follows: minimap2 2.24 (requirement.yaml:12), the caller of ksw2 behind FocalSV's
// `minimap2 -a -x asm5 --cs -r2k` (focalsv/4_sv_calling/Dippav/DipPAV_variant_call.py:103): sketch.c (minimizers),
// chain.c (mm_chain_dp: score min(d, k) minus a gap cost 0.01*k*|dd| + 0.5*log2|dd|), align.c (mm_align1: fills between
// chain anchors at least min_ksw_len apart).  That code is NOT in /root/reference, so this is a restatement of the
// published algorithm, PARITY UNPINNED: same strand only, one chain, global ends (the hook's windows span the contig),
// no z-drop re-runs.  Every task it emits is an ordinary fsv_task, so the DP itself stays bit-exact ksw2.
// Host only: no device work, usable without a GPU.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/focalsv_cuda.h"

namespace {

struct Mz { uint64_t h; int32_t pos; };      // hash of the k-mer ENDING at pos (inclusive)

inline uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

// (w, k)-minimizers of the forward strand: the k-mer with the smallest hash in every window of w consecutive k-mers
// (leftmost on ties); k-mers holding a wildcard (code > 3) do not exist.
void sketch(const uint8_t* s, int32_t n, int k, int w, std::vector<Mz>& out)
{
    out.clear();
    if (n < k) return;
    const uint64_t mask = k < 32 ? (1ull << (2 * k)) - 1 : ~0ull;
    std::vector<Mz> ring((size_t)w);
    uint64_t kmer = 0;
    int valid = 0, filled = 0, last_pos = -1;
    for (int32_t i = 0; i < n; ++i) {
        if (s[i] > 3) { valid = 0; filled = 0; continue; }
        kmer = ((kmer << 2) | s[i]) & mask;
        if (++valid < k) continue;
        ring[(size_t)(filled % w)] = Mz{mix64(kmer), i};
        ++filled;
        if (filled < w) continue;
        int best = 0;
        for (int j = 1; j < w; ++j)
            if (ring[(size_t)j].h < ring[(size_t)best].h || (ring[(size_t)j].h == ring[(size_t)best].h && ring[(size_t)j].pos < ring[(size_t)best].pos)) best = j;
        if (ring[(size_t)best].pos != last_pos) { out.push_back(ring[(size_t)best]); last_pos = ring[(size_t)best].pos; }
    }
}

struct Anchor { int32_t t, q; };      // last base of the shared k-mer in the target / in the query

}  // namespace

extern "C" int fsv_chain_pieces(const uint8_t* query, int32_t qlen, const uint8_t* target, int32_t tlen,
                                int k, int w, int max_occ, int max_gap, int min_fill,
                                fsv_piece* pieces, size_t cap, size_t* n_pieces, int32_t* chain_score, int32_t* n_anchors)
{
    if (!query || !target || qlen < 0 || tlen < 0 || k < 4 || k > 31 || w < 1 || w > 256 || max_occ < 1 || max_gap < 1 || min_fill < 1 || !n_pieces)
        return FSV_ERR_INVALID;
    std::vector<Mz> mq, mt;
    sketch(query, qlen, k, w, mq);
    sketch(target, tlen, k, w, mt);
    std::sort(mt.begin(), mt.end(), [](const Mz& a, const Mz& b) { return a.h < b.h || (a.h == b.h && a.pos < b.pos); });
    std::vector<Anchor> a;
    for (const Mz& m : mq) {
        auto lo = std::lower_bound(mt.begin(), mt.end(), m.h, [](const Mz& x, uint64_t h) { return x.h < h; });
        auto hi = lo;
        while (hi != mt.end() && hi->h == m.h) ++hi;
        if (hi - lo > max_occ) continue;                      // repetitive seed (mid_occ filter)
        for (auto it = lo; it != hi; ++it) a.push_back(Anchor{it->pos, m.pos});
    }
    std::sort(a.begin(), a.end(), [](const Anchor& x, const Anchor& y) { return x.t < y.t || (x.t == y.t && x.q < y.q); });
    const int n = (int)a.size();
    // chaining (mm_chain_dp): f[i] = max(k, max_j f[j] + min(dq, dt, k) - gap(|dq - dt|)) over predecessors with
    // 0 < dq, dt <= max_gap, looking back at most 64 anchors past the last improvement
    std::vector<float> f((size_t)n);
    std::vector<int32_t> p((size_t)n, -1);
    int best = -1;
    for (int i = 0; i < n; ++i) {
        float fi = (float)k;
        int pi = -1, since = 0;
        for (int j = i - 1; j >= 0 && since < 64; --j) {
            const int dt = a[(size_t)i].t - a[(size_t)j].t, dq = a[(size_t)i].q - a[(size_t)j].q;
            if (dt > max_gap) break;
            ++since;
            if (dt <= 0 || dq <= 0 || dq > max_gap) continue;
            const int dd = dt > dq ? dt - dq : dq - dt;
            float sc = (float)std::min(std::min(dt, dq), k);
            if (dd) sc -= 0.01f * (float)k * (float)dd + 0.5f * log2f((float)dd);
            if (f[(size_t)j] + sc > fi) { fi = f[(size_t)j] + sc; pi = j; since = 0; }
        }
        f[(size_t)i] = fi; p[(size_t)i] = pi;
        if (best < 0 || fi > f[(size_t)best]) best = i;
    }
    std::vector<Anchor> chain;
    for (int i = best; i >= 0; i = p[(size_t)i]) chain.push_back(a[(size_t)i]);
    std::reverse(chain.begin(), chain.end());
    if (chain_score) *chain_score = best >= 0 ? (int32_t)f[(size_t)best] : 0;
    if (n_anchors) *n_anchors = (int32_t)chain.size();
    // decomposition (mm_align1): fills from the END of the last cut anchor to the END of the next anchor that lies at
    // least min_fill further (anchors in between are not trusted, they lie inside the fill); global ends
    std::vector<fsv_piece> out;
    int32_t qs = 0, ts = 0;
    auto push = [&](int32_t qe, int32_t te) {
        if (qe == qs && te == ts) return;
        out.push_back(fsv_piece{qs, qe, ts, te});
        qs = qe; ts = te;
    };
    if (!chain.empty()) {
        const int32_t q0 = chain[0].q - k + 1, t0 = chain[0].t - k + 1;      // start of the first shared k-mer
        push(q0, t0);
        for (size_t i = 0; i < chain.size(); ++i) {
            const int32_t qe = chain[i].q + 1, te = chain[i].t + 1;
            if (qe <= qs || te <= ts) continue;                               // overlaps what is already covered
            if (i + 1 == chain.size() || qe - qs >= min_fill || te - ts >= min_fill) push(qe, te);
        }
    }
    push(qlen, tlen);
    *n_pieces = out.size();
    if (out.size() > cap || (!pieces && !out.empty())) return FSV_ERR_CIGAR_CAP;      // *n_pieces = entries needed
    if (!out.empty()) memcpy(pieces, out.data(), out.size() * sizeof(fsv_piece));
    return FSV_OK;
}

// Stitch the CIGARs of one pair's pieces (in order) into one: a piece with a task contributes that task's CIGAR, a piece
// with an empty side a pure gap; neighbouring operations of the same kind are merged (BAM words, len << 4 | op).
extern "C" int fsv_stitch_cigars(const fsv_piece* pieces, const int32_t* task_of, size_t n_pieces,
                                 const fsv_result* res, const uint32_t* cigar_arena,
                                 uint32_t* out, size_t cap, size_t* n_out)
{
    if ((!pieces || !task_of) && n_pieces) return FSV_ERR_INVALID;
    if (!n_out) return FSV_ERR_INVALID;
    size_t n = 0;
    auto push = [&](uint32_t op, uint64_t len) {
        if (!len) return;
        if (n && n <= cap && (out[n - 1] & 0xfu) == op && (uint64_t)(out[n - 1] >> 4) + len < (1u << 28)) { out[n - 1] += (uint32_t)len << 4; return; }
        if (n < cap) out[n] = (uint32_t)len << 4 | op;
        ++n;
    };
    for (size_t i = 0; i < n_pieces; ++i) {
        const int32_t dq = pieces[i].q_end - pieces[i].q_beg, dt = pieces[i].t_end - pieces[i].t_beg;
        if (task_of[i] >= 0) {
            if (!res || !cigar_arena) return FSV_ERR_INVALID;
            const fsv_result& r = res[task_of[i]];
            for (int32_t k = 0; k < r.n_cigar; ++k) push(cigar_arena[r.cigar_off + k] & 0xfu, cigar_arena[r.cigar_off + k] >> 4);
        } else if (dq > 0) push(1u, (uint64_t)dq);
        else if (dt > 0) push(2u, (uint64_t)dt);
    }
    *n_out = n;
    return n > cap ? FSV_ERR_CIGAR_CAP : FSV_OK;
}
