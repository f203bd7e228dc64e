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
    // sliding-window minimum over the last w k-mers as a monotonic queue (O(n) instead of O(n w)): dq[head .. tail) holds the candidates in
    // order of position with strictly increasing hash from back to front removed, so dq[head] is the smallest hash of the window, leftmost on ties
    std::vector<Mz> dq((size_t)w + 1);
    int head = 0, tail = 0;          // ring indices into dq, at most w entries
    auto at = [&](int i) -> Mz& { return dq[(size_t)(i % (w + 1))]; };
    uint64_t kmer = 0;
    int valid = 0, filled = 0, last_pos = -1;
    for (int32_t i = 0; i < n; ++i) {
        if (s[i] > 3) { valid = 0; filled = 0; head = tail = 0; continue; }
        kmer = ((kmer << 2) | s[i]) & mask;
        if (++valid < k) continue;
        const Mz e{mix64(kmer), i};
        while (tail > head && at(tail - 1).h > e.h) --tail;      // a later k-mer wins only with a strictly smaller hash
        at(tail) = e; ++tail;
        ++filled;
        if (at(head).pos <= i - w) ++head;                       // left the window of the last w k-mers
        if (filled < w) continue;
        const Mz& b = at(head);
        if (b.pos != last_pos) { out.push_back(b); last_pos = b.pos; }
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

// ---------------------------------------------------------------------------------------------------------------------
// Second version (row f2): both strands, several chains per pair, extension ends.
//
// What `minimap2 -a -x <preset>` does with one contig before and around its ksw2 calls, restated from the published
// algorithm (minimap2 2.24 sketch.c / chain.c / align.c are NOT in the reference tree: PARITY UNPINNED):
//   * canonical (w,k)-minimizers: a k-mer and its reverse complement hash alike and carry a strand bit (palindromes skipped);
//   * anchors of equal and of opposite strand are chained separately (the query of the reverse chains is the reverse
//     complement); chains are peeled off best-first (mm_chain_backtrack: a chain stops where it meets an anchor already used);
//   * the best chain is the primary; a chain whose query interval overlaps accepted ones by less than half of its own length is
//     a supplementary alignment (the split alignments the reference reads through SA tags / consecutive records,
//     extract_contig_signature_CCS.py:268-327, svim-asm SVIM_COLLECT.py:8-54), the others (secondaries) are dropped;
//   * per chain: the core between the first and the last anchor is cut into global fills at anchors >= min_fill apart (as
//     fsv_chain_pieces does), and the two ends are OFFERED for extension: lq / rq query bases and lt / rt target bases
//     (mm_align1: the query end plus the longest gap its score could pay for, capped by max_gap and the sequence ends).
// Host only.
namespace {

struct Mz2 { uint64_t h; int32_t pos; int32_t strand; };

// canonical minimizers; pos = last base of the k-mer on the FORWARD sequence, strand = 1 when the reverse complement is the smaller one
void sketch2(const uint8_t* s, int32_t n, int k, int w, std::vector<Mz2>& out)
{
    out.clear();
    if (n < k) return;
    const uint64_t mask = k < 32 ? (1ull << (2 * k)) - 1 : ~0ull;
    const int shift = 2 * (k - 1);
    std::vector<Mz2> dq((size_t)w + 1);                          // monotonic queue, as in sketch()
    int head = 0, tail = 0;
    auto at = [&](int i) -> Mz2& { return dq[(size_t)(i % (w + 1))]; };
    uint64_t fw = 0, rv = 0;
    int valid = 0, filled = 0, last_pos = -1;
    for (int32_t i = 0; i < n; ++i) {
        if (s[i] > 3) { valid = 0; filled = 0; head = tail = 0; continue; }
        fw = ((fw << 2) | s[i]) & mask;
        rv = (rv >> 2) | ((uint64_t)(3 - s[i]) << shift);
        if (++valid < k) continue;
        // a palindromic k-mer has no strand: it takes a slot in the window (so that windows stay w k-mers wide) but never wins
        const bool pal = fw == rv;
        const Mz2 e{pal ? ~0ull : mix64(fw < rv ? fw : rv), i, fw < rv ? 0 : 1};
        while (tail > head && at(tail - 1).h > e.h) --tail;
        at(tail) = e; ++tail;
        ++filled;
        if (at(head).pos <= i - w) ++head;
        if (filled < w) continue;
        const Mz2& b = at(head);
        if (b.h != ~0ull && b.pos != last_pos) { out.push_back(b); last_pos = b.pos; }
    }
}

struct Chain2 { int strand; float score; std::vector<Anchor> a; int32_t qb, qe; /* query interval on the FORWARD query */ };

// chaining DP of fsv_chain_pieces over one strand's anchors, then best-first peeling
void chain_strand(std::vector<Anchor>& a, int strand, int32_t qlen, int k, int max_gap, float min_score, int min_cnt, std::vector<Chain2>& out)
{
    std::sort(a.begin(), a.end(), [](const Anchor& x, const Anchor& y) { return x.t < y.t || (x.t == y.t && x.q < y.q); });
    const int n = (int)a.size();
    if (!n) return;
    std::vector<float> f((size_t)n);
    std::vector<int32_t> p((size_t)n, -1);
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
    }
    std::vector<int32_t> order((size_t)n);
    for (int i = 0; i < n; ++i) order[(size_t)i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return f[(size_t)x] > f[(size_t)y]; });
    std::vector<uint8_t> used((size_t)n, 0);
    for (int32_t end : order) {
        if (used[(size_t)end] || f[(size_t)end] < min_score) continue;
        std::vector<Anchor> c;
        int i = end;
        for (; i >= 0 && !used[(size_t)i]; i = p[(size_t)i]) c.push_back(a[(size_t)i]);
        const float sc = f[(size_t)end] - (i >= 0 ? f[(size_t)i] : 0.f);
        if (sc < min_score || (int)c.size() < min_cnt) continue;      // (the anchors stay free for a better start)
        for (int j = end, cnt = (int)c.size(); cnt > 0; j = p[(size_t)j], --cnt) used[(size_t)j] = 1;
        std::reverse(c.begin(), c.end());
        Chain2 ch; ch.strand = strand; ch.score = sc;
        const int32_t q0 = c.front().q - k + 1, q1 = c.back().q + 1;                  // on the strand-adjusted query
        ch.qb = strand ? qlen - q1 : q0; ch.qe = strand ? qlen - q0 : q1;
        ch.a.swap(c);
        out.push_back(std::move(ch));
    }
}

}  // namespace

extern "C" int fsv_chain_pair(const uint8_t* query, int32_t qlen, const uint8_t* target, int32_t tlen, const fsv_chain_opts* o,
                              fsv_chain* chains, size_t chain_cap, size_t* n_chains,
                              fsv_piece* pieces, size_t piece_cap, size_t* n_pieces)
{
    if (!query || !target || !o || qlen < 0 || tlen < 0 || !n_chains || !n_pieces) return FSV_ERR_INVALID;
    const int k = o->k, w = o->w;
    if (k < 4 || k > 31 || w < 1 || w > 256 || o->max_occ < 1 || o->max_gap < 1 || o->min_fill < 1 || o->max_chains < 1 || o->e < 1) return FSV_ERR_INVALID;
    std::vector<Mz2> mq, mt;
    sketch2(query, qlen, k, w, mq);
    sketch2(target, tlen, k, w, mt);
    std::sort(mt.begin(), mt.end(), [](const Mz2& a, const Mz2& b) { return a.h < b.h || (a.h == b.h && a.pos < b.pos); });
    std::vector<Anchor> fwd, rev;
    for (const Mz2& m : mq) {
        auto lo = std::lower_bound(mt.begin(), mt.end(), m.h, [](const Mz2& x, uint64_t h) { return x.h < h; });
        auto hi = lo;
        while (hi != mt.end() && hi->h == m.h) ++hi;
        if (hi - lo > o->max_occ) continue;
        for (auto it = lo; it != hi; ++it) {
            if (it->strand == m.strand) fwd.push_back(Anchor{it->pos, m.pos});
            else rev.push_back(Anchor{it->pos, qlen - 1 - (m.pos - k + 1)});        // end of the k-mer on the reverse-complemented query
        }
    }
    std::vector<Chain2> all;
    chain_strand(fwd, 0, qlen, k, o->max_gap, (float)o->min_chain_score, o->min_anchors, all);
    chain_strand(rev, 1, qlen, k, o->max_gap, (float)o->min_chain_score, o->min_anchors, all);
    std::stable_sort(all.begin(), all.end(), [](const Chain2& x, const Chain2& y) { return x.score > y.score; });
    // primary / supplementary / secondary by query overlap with the chains accepted so far.  One case minimap2 resolves later,
    // by the z-drop inside a gap fill, is resolved here on the anchors: a chain that sits in a HOLE of an accepted chain
    // (no anchor of that chain inside its query interval: an inverted or replaced segment the longer chain jumped over) is not
    // a secondary of it; the longer chain is split at the hole and all three become alignments of their own.
    auto fwd_interval = [&](const Chain2& c, const Anchor& an, int32_t& b, int32_t& e) {      // k-mer of an anchor on the FORWARD query
        const int32_t q0 = an.q - k + 1, q1 = an.q + 1;
        b = c.strand ? qlen - q1 : q0; e = c.strand ? qlen - q0 : q1;
    };
    std::vector<Chain2> kept;
    std::vector<float> sub;
    for (size_t i = 0; i < all.size(); ++i) {
        Chain2 c = all[i];
        int parent = -1, hole_of = -1;
        size_t hole_at = 0;
        for (size_t j = 0; j < kept.size() && parent < 0 && hole_of < 0; ++j) {
            const Chain2& P = kept[j];
            const int32_t ov = std::min(c.qe, P.qe) - std::max(c.qb, P.qb);
            if (ov <= 0 || 2 * (int64_t)ov < std::min<int64_t>(c.qe - c.qb, P.qe - P.qb)) continue;
            // anchors of P inside c's interval?
            size_t first_after = P.a.size(), n_inside = 0;
            for (size_t x = 0; x < P.a.size(); ++x) {
                int32_t b, e; fwd_interval(P, P.a[x], b, e);
                if (e > c.qb + k && b < c.qe - k) ++n_inside;
            }
            if (n_inside == 0 && c.qb >= P.qb && c.qe <= P.qe) {
                // the anchors of P are ordered along ITS query; find the split point: first anchor past the hole on P's own strand
                for (size_t x = 0; x < P.a.size(); ++x) {
                    int32_t b, e; fwd_interval(P, P.a[x], b, e);
                    const bool past = P.strand ? e <= c.qb + k : b >= c.qe - k;
                    if (past) { first_after = x; break; }
                }
                if (first_after > 0 && first_after < P.a.size() && (int)first_after >= o->min_anchors && (int)(P.a.size() - first_after) >= o->min_anchors) { hole_of = (int)j; hole_at = first_after; }
                else parent = (int)j;
            } else parent = (int)j;
        }
        if (parent >= 0) { sub[(size_t)parent] = std::max(sub[(size_t)parent], c.score); continue; }      // a secondary: only its score is remembered (mapq)
        if (hole_of >= 0 && (int)kept.size() + 2 <= o->max_chains) {
            Chain2 P = kept[(size_t)hole_of];
            Chain2 A, B;
            A.strand = B.strand = P.strand;
            A.a.assign(P.a.begin(), P.a.begin() + (long)hole_at); B.a.assign(P.a.begin() + (long)hole_at, P.a.end());
            A.score = P.score * (float)A.a.size() / (float)P.a.size(); B.score = P.score - A.score;
            for (Chain2* h : {&A, &B}) {
                const int32_t q0 = h->a.front().q - k + 1, q1 = h->a.back().q + 1;
                h->qb = h->strand ? qlen - q1 : q0; h->qe = h->strand ? qlen - q0 : q1;
            }
            const bool a_first = A.score >= B.score;
            kept[(size_t)hole_of] = a_first ? A : B;
            kept.push_back(a_first ? B : A); sub.push_back(0.f);
            kept.push_back(c); sub.push_back(0.f);
            continue;
        }
        if (hole_of >= 0) { sub[(size_t)hole_of] = std::max(sub[(size_t)hole_of], c.score); continue; }
        if ((int)kept.size() < o->max_chains) { kept.push_back(c); sub.push_back(0.f); }
    }
    std::vector<const Chain2*> keep;
    for (const Chain2& c : kept) keep.push_back(&c);
    std::vector<fsv_chain> oc;
    std::vector<fsv_piece> op;
    for (size_t ci = 0; ci < keep.size(); ++ci) {
        const Chain2& c = *keep[ci];
        fsv_chain r{};
        r.strand = c.strand; r.score = (int32_t)c.score; r.n_anchors = (int32_t)c.a.size(); r.sub_score = (int32_t)sub[ci];
        r.q_beg = c.a.front().q - k + 1; r.t_beg = c.a.front().t - k + 1; r.q_end = c.a.back().q + 1; r.t_end = c.a.back().t + 1;
        r.piece_off = (int32_t)op.size();
        int32_t qs = r.q_beg, ts = r.t_beg;
        for (size_t i = 0; i < c.a.size(); ++i) {
            const int32_t qe = c.a[i].q + 1, te = c.a[i].t + 1;
            if (qe <= qs || te <= ts) continue;
            if (i + 1 == c.a.size() || qe - qs >= o->min_fill || te - ts >= o->min_fill) { op.push_back(fsv_piece{qs, qe, ts, te}); qs = qe; ts = te; }
        }
        r.q_end = qs; r.t_end = ts;                       // (a trailing anchor that overlapped what was covered is not part of the core)
        r.n_pieces = (int32_t)op.size() - r.piece_off;
        // extension offers (mm_align1): the query end, and on the target that many bases plus the longest gap the end could pay for
        auto offer = [&](int32_t qleft, int32_t tleft, int32_t& oq, int32_t& ot) {
            int64_t l = std::min<int64_t>(qleft, o->max_gap);
            oq = (int32_t)l;
            const int64_t pay = l * o->a + o->end_bonus;
            l += pay > o->q ? (pay - o->q) / o->e : 0;
            l = std::min<int64_t>(std::min<int64_t>(l, o->max_gap), tleft);
            ot = (int32_t)l;
        };
        offer(r.q_beg, r.t_beg, r.lq, r.lt);
        offer(qlen - r.q_end, tlen - r.t_end, r.rq, r.rt);
        oc.push_back(r);
    }
    *n_chains = oc.size(); *n_pieces = op.size();
    if (oc.size() > chain_cap || op.size() > piece_cap || (!chains && !oc.empty()) || (!pieces && !op.empty())) return FSV_ERR_CIGAR_CAP;
    if (!oc.empty()) memcpy(chains, oc.data(), oc.size() * sizeof(fsv_chain));
    if (!op.empty()) memcpy(pieces, op.data(), op.size() * sizeof(fsv_piece));
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
