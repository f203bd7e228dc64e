"""Pipeline hook: (contig, reference window) pairs in, alignment records and per-contig SV signatures out.

This is the host-side mirror of the step FocalSV performs around its minimap2 call
(focalsv/4_sv_calling/Dippav/DipPAV_variant_call.py:97-137): align every haplotype contig of a
region against that region's reference window, then walk each CIGAR for DEL/INS signatures
(extract_contig_signature_CCS.py:14-127, `extract_sig_from_cigar`).  The alignment itself is
`Aligner.align_batch` (libfocalsv_cuda); this module only shapes inputs and outputs so that the
reference's consumers see what they read from pysam today: reference_name, pos, cigar tuples,
qname, is_reverse, mapq.

Next-row scope: the signature extraction below is the reference's per-read rule set
restated in a few lines of Python; clustering across contigs / genotype pairing stay in the
reference's own code, which consumes these records unchanged.
"""
from collections import namedtuple

import numpy as np

from . import _abi
from .api import make_tasks, task_cigar
from .presets import PRESETS, ksw_band, scoring_for

AlignedContig = namedtuple("AlignedContig", "qname reference_name pos reference_end cigar is_reverse mapq query_length score zdropped")
Signature = namedtuple("Signature", "chrom svtype pos svlen qname read_start read_end strand source mapq")

_CODE = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate("ACGT"):
    _CODE[ord(_c)] = _i
    _CODE[ord(_c.lower())] = _i


def encode(seq):
    """ASCII bases -> codes 0..3, everything else 4 (Correct.cpp:7676-7685 convention)."""
    if isinstance(seq, np.ndarray) and seq.dtype == np.uint8 and (seq.size == 0 or seq.max() <= 4):
        return seq
    if isinstance(seq, str):
        seq = seq.encode()
    return _CODE[np.frombuffer(bytes(seq), dtype=np.uint8)]


def cigar_tuples(words):
    """BAM-encoded words -> pysam-style [(op, len), ...]."""
    return [(int(w) & 0xF, int(w) >> 4) for w in words]


def realign_regions(aligner, windows, contigs, preset="asm5", bw=2000, flag=0, zdrop=None):
    """Align contig i against window i (same region) on the GPU.

    windows: list of (chrom, start, sequence); contigs: list of (qname, sequence).
    Scoring / z-drop are the preset's (presets.py); the band is minimap2's bw*1.5+1 for `-r{bw}`
    (DipPAV_variant_call.py:103 passes -r2k); `zdrop` overrides the preset's.  Returns one AlignedContig per pair."""
    p = PRESETS[preset]
    sc = scoring_for(preset)
    q = [encode(s) for _, s in contigs]
    t = [encode(s) for _, _, s in windows]
    tasks = make_tasks([len(x) for x in q], [len(x) for x in t], ksw_band(bw), p.zdrop if zdrop is None else zdrop, 0, flag)
    qa = np.concatenate(q) if q else np.zeros(0, np.uint8)
    ta = np.concatenate(t) if t else np.zeros(0, np.uint8)
    res, arena = aligner.align_batch(sc, qa, ta, tasks)
    return records_from_results(windows, contigs, res, arena)


class MultiAligner(object):
    """Several aligners (one per GPU of this host) behind one align_batch: SURVEY 8e's in-process form of the multi-GPU path.
    Tasks are independent (ksw_extz2_sse has no global state), so a batch is cut into one length-balanced bin per device
    (fsv_lpt_bins: longest-processing-time-first over the estimated cells), every device runs its bin on its own host thread
    (the C call releases the GIL), and the results come back IN THE CALLER'S TASK ORDER with the CIGARs re-packed into one
    arena.  No collective: this is a host-side gather.  `aligners` is a list of anything with align_batch (api.Aligner(dev))."""

    def __init__(self, aligners):
        if not aligners:
            raise ValueError("MultiAligner needs at least one aligner")
        self.aligners = list(aligners)

    @classmethod
    def on_devices(cls, devices):
        from .api import Aligner
        return cls([Aligner(d) for d in devices])

    def close(self):
        for a in self.aligners:
            if hasattr(a, "close"):
                a.close()

    def align_batch(self, sc, qarena, tarena, tasks):
        from concurrent.futures import ThreadPoolExecutor
        from .api import lpt_bins
        tasks = np.ascontiguousarray(tasks, dtype=_abi.TASK_DTYPE)
        n, nd = len(tasks), len(self.aligners)
        if nd == 1 or n == 0:
            return self.aligners[0].align_batch(sc, qarena, tarena, tasks)
        bins = lpt_bins(tasks, nd)
        idx = [np.flatnonzero(bins == d) for d in range(nd)]
        with ThreadPoolExecutor(max_workers=nd) as ex:
            parts = list(ex.map(lambda d: self.aligners[d].align_batch(sc, qarena, tarena, tasks[idx[d]]) if len(idx[d]) else None, range(nd)))
        res = np.zeros(n, dtype=_abi.RESULT_DTYPE)
        for d in range(nd):
            if parts[d] is not None:
                res[idx[d]] = parts[d][0]
        # one arena in task order
        off = np.concatenate([[0], np.cumsum(res["n_cigar"].astype(np.int64))])
        arena = np.zeros(int(off[-1]), dtype=np.uint32)
        for d in range(nd):
            if parts[d] is None:
                continue
            r, a = parts[d]
            for k, i in enumerate(idx[d]):
                o, m = int(r[k]["cigar_off"]), int(r[k]["n_cigar"])
                arena[off[i]:off[i] + m] = a[o:o + m]
        res["cigar_off"] = off[:-1]
        return res, arena


def realign_regions_abi(aligner, ref_codes, regions, contigs, preset="asm5", bw=2000, flag=0):
    """The same step through the library's own Level-1 entry point (fsv_realign_regions): the reference is passed
    once, regions are (chrom, start, end) into it and are not copied on the host.  Returns AlignedContig records."""
    q = [encode(s) for _, s in contigs]
    lens = np.array([len(x) for x in q], dtype=np.int32)
    offs = np.concatenate([[0], np.cumsum(lens.astype(np.int64))[:-1]]) if len(q) else np.zeros(0, np.int64)
    rec, arena = aligner.realign_regions(encode(ref_codes), [s for _, s, _ in regions], [e for _, _, e in regions],
                                         np.concatenate(q) if q else np.zeros(0, np.uint8), offs, lens, preset, bw, flag)
    out = []
    for r, (chrom, _, _), (qname, _) in zip(rec, regions, contigs):
        cig = cigar_tuples(arena[int(r["cigar_off"]):int(r["cigar_off"]) + int(r["n_cigar"])])
        q_used = sum(n for op, n in cig if op in (0, 1, 4))
        if q_used < int(r["query_length"]):                   # z-dropped: soft clip for the unaligned tail (see records_from_results)
            cig = cig + [(4, int(r["query_length"]) - q_used)]
        out.append(AlignedContig(qname, chrom, int(r["pos"]), int(r["ref_end"]), cig, bool(r["is_reverse"]), int(r["mapq"]),
                                 int(r["query_length"]), int(r["score"]), bool(r["zdropped"])))
    return out


# minimap2 2.24 preset seeding parameters (-k, -w) [UNVERIFIED-EXT, SURVEY appendix B]
SEEDING = {"asm5": (19, 19), "asm10": (19, 19), "map-hifi": (19, 19), "map-pb": (19, 10), "map-ont": (15, 10)}


def realign_regions_chained(aligner, windows, contigs, preset="asm5", bw=2000, min_fill=200, max_occ=50):
    """Row f2: the same (contig, window) pairs as realign_regions, but decomposed the way minimap2 decomposes them
    before it calls ksw2 (api.chain_pieces: minimizers, chaining, cuts at anchors >= min_fill apart): every piece is a
    small global DP task, all pieces of all pairs go to the GPU as ONE batch, and the piece CIGARs are stitched.
    `aligner` is anything with align_batch(sc, qarena, tarena, tasks) -> (results, cigar_arena).  The result is a valid
    global alignment of each pair; it equals minimap2's only as far as the restated seeding does (parity unpinned)."""
    from .api import chain_pieces
    if preset not in SEEDING:
        raise ValueError("realign_regions_chained: preset %r has no seeding parameters (minimap2 presets only: %s)" % (preset, ", ".join(sorted(SEEDING))))
    p = PRESETS[preset]
    sc = scoring_for(preset)
    k, w = SEEDING[preset]
    q = [encode(s) for _, s in contigs]
    t = [encode(s) for _, _, s in windows]
    q_off = np.concatenate([[0], np.cumsum([len(x) for x in q])]).astype(np.int64)
    t_off = np.concatenate([[0], np.cumsum([len(x) for x in t])]).astype(np.int64)
    per_pair, tasks_l, n_tasks = [], [], 0
    # seeding + chaining is host work per pair and the C call releases the GIL: a few threads, pairs in order
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(16, max(1, len(q)))) as ex:
        chained = list(ex.map(lambda i: chain_pieces(q[i], t[i], k, w, max_occ, p.bw_long, min_fill)[0], range(len(q))))
    for i in range(len(q)):
        pcs = chained[i]
        dq = (pcs["q_end"] - pcs["q_beg"]).astype(np.int64); dt = (pcs["t_end"] - pcs["t_beg"]).astype(np.int64)
        has = (dq > 0) & (dt > 0)
        task_of = np.where(has, n_tasks + np.cumsum(has) - 1, -1).astype(np.int32)
        tk = np.zeros(int(has.sum()), dtype=_abi.TASK_DTYPE)
        tk["q_off"] = q_off[i] + pcs["q_beg"][has]; tk["t_off"] = t_off[i] + pcs["t_beg"][has]
        tk["qlen"] = dq[has]; tk["tlen"] = dt[has]
        tk["w"] = np.abs(dq - dt)[has] + min(ksw_band(bw), 200)
        tk["zdrop"] = -1                                      # global fills: no z-drop (minimap2's zdrop_inv re-runs are not restated)
        tasks_l.append(tk); n_tasks += len(tk)
        per_pair.append((pcs, task_of, dq, dt, has))
    tasks = np.concatenate(tasks_l) if tasks_l else np.zeros(0, dtype=_abi.TASK_DTYPE)
    res, arena = aligner.align_batch(sc, np.concatenate(q) if q else np.zeros(0, np.uint8),
                                     np.concatenate(t) if t else np.zeros(0, np.uint8), tasks)
    from .api import stitch_cigars
    if len(res) and (res["status"] != 0).any():
        raise ValueError("realign_regions_chained: %d piece tasks were reset (scoring outside ksw2's int8 range)" % int((res["status"] != 0).sum()))
    if len(res) and (res["zdropped"] != 0).any():
        # global fills run without z-drop; a piece can still stop early when |dq - dt| exceeds its band (ksw2_extz2_sse.c:111-114)
        raise ValueError("realign_regions_chained: %d pieces ran out of band" % int((res["zdropped"] != 0).sum()))
    cig, score = [], []
    for pcs, task_of, dq, dt, has in per_pair:
        cig.append(cigar_tuples(stitch_cigars(pcs, task_of, res, arena)))
        gaps = np.where(has, 0, np.maximum(dq, dt))           # pieces with an empty side: one gap of the other side's length
        gcost = np.minimum(p.q + gaps * p.e, p.q2 + gaps * p.e2) if p.q2 >= 0 else p.q + gaps * p.e
        score.append(int(res["score"][task_of[has]].astype(np.int64).sum()) - int(gcost[gaps > 0].sum()))
    out = []
    for i, ((chrom, start, _), (qname, _)) in enumerate(zip(windows, contigs)):
        ref_span = sum(n for op, n in cig[i] if op in (0, 2))
        out.append(AlignedContig(qname, chrom, int(start), int(start) + ref_span, cig[i], False, 60, len(q[i]), score[i], False))
    return out


def _revcomp_codes(x):
    """reverse complement of a code array (0..3 -> 3 - code, wildcards stay)"""
    r = x[::-1].copy()
    m = r < 4
    r[m] = 3 - r[m]
    return r


def _merge_cigar(parts):
    out = []
    for part in parts:
        for op, n in part:
            if n <= 0:
                continue
            if out and out[-1][0] == op:
                out[-1] = (op, out[-1][1] + n)
            else:
                out.append((op, n))
    return out


def map_contigs(aligner, windows, contigs, preset="asm5", bw=2000, max_chains=8, min_fill=200, max_occ=50):
    """Row f2, second version: what `minimap2 -a -x <preset> -r<bw>` returns for every contig against its window, restated
    (PARITY UNPINNED): strand selection, one primary and any number of supplementary alignments per contig (split
    at inversions / replaced segments), ends found by extension with the preset's z-drop instead of being forced to the
    window's corners, clips, mapping quality, NM / AS / SA tags.  All DP work of all contigs is ONE GPU batch: the global fills
    between chain anchors (fsv_chain_pair) plus two extension tasks per alignment (left: reversed sequences with
    KSW_EZ_EXTZ_ONLY | KSW_EZ_RIGHT | KSW_EZ_REV_CIGAR; right: KSW_EZ_EXTZ_ONLY), every one an ordinary bit-exact ksw2 task.
    Returns dropin.SamRecord objects (primary first per contig, then its supplementary alignments)."""
    from concurrent.futures import ThreadPoolExecutor
    from .api import chain_pair, stitch_cigars
    from .dropin import SamRecord, edit_distance_tag
    if preset not in SEEDING:
        raise ValueError("map_contigs: preset %r has no seeding parameters (minimap2 presets only: %s)" % (preset, ", ".join(sorted(SEEDING))))
    p = PRESETS[preset]
    sc = scoring_for(preset)
    k, w = SEEDING[preset]
    end_bonus = max(p.end_bonus, 0)
    opts = _abi.ChainOpts(k, w, max_occ, p.bw_long, min_fill, max_chains, 40, 3, p.a, p.q, p.e, end_bonus)
    qf = [encode(s_) for _, s_ in contigs]
    tt = [encode(s_) for _, _, s_ in windows]
    with ThreadPoolExecutor(max_workers=min(16, max(1, len(qf)))) as ex:
        chained = list(ex.map(lambda i: chain_pair(qf[i], tt[i], opts), range(len(qf))))
    band = ksw_band(bw)
    qparts, tparts, rows = [], [], []          # task sequences (copies) and one row per task: (qlen, tlen, w, zdrop, end_bonus, flag)
    plan = []                                   # per pair: list of (chain, pieces, query codes on the chain's strand, task ids)

    def add_task(qseq, tseq, w_, zdrop, eb, flag):
        qparts.append(qseq); tparts.append(tseq)
        rows.append((len(qseq), len(tseq), w_, zdrop, eb, flag))
        return len(rows) - 1

    for i, (chains, pieces) in enumerate(chained):
        per = []
        qr = None
        for c in chains:
            if int(c["strand"]) and qr is None:
                qr = _revcomp_codes(qf[i])
            qs = qr if int(c["strand"]) else qf[i]
            pcs = pieces[int(c["piece_off"]):int(c["piece_off"]) + int(c["n_pieces"])]
            ids = []
            for pc in pcs:
                dq, dt = int(pc["q_end"] - pc["q_beg"]), int(pc["t_end"] - pc["t_beg"])
                ids.append(add_task(qs[pc["q_beg"]:pc["q_end"]], tt[i][pc["t_beg"]:pc["t_end"]], abs(dq - dt) + min(band, 200), -1, 0, 0)
                           if dq > 0 and dt > 0 else -1)
            lq, lt, rq, rt = int(c["lq"]), int(c["lt"]), int(c["rq"]), int(c["rt"])
            qb, tb, qe, te = int(c["q_beg"]), int(c["t_beg"]), int(c["q_end"]), int(c["t_end"])
            left = add_task(qs[qb - lq:qb][::-1], tt[i][tb - lt:tb][::-1], band, p.zdrop, end_bonus,
                            _abi.EZ_EXTZ_ONLY | _abi.EZ_RIGHT | _abi.EZ_REV_CIGAR) if lq > 0 and lt > 0 else -1
            right = add_task(qs[qe:qe + rq], tt[i][te:te + rt], band, p.zdrop, end_bonus, _abi.EZ_EXTZ_ONLY) if rq > 0 and rt > 0 else -1
            per.append((c, pcs, qs, ids, left, right))
        plan.append(per)
    tasks = np.zeros(len(rows), dtype=_abi.TASK_DTYPE)
    if rows:
        arr = np.array(rows, dtype=np.int64)
        tasks["qlen"], tasks["tlen"], tasks["w"], tasks["zdrop"], tasks["end_bonus"], tasks["flag"] = arr.T
        tasks["q_off"] = np.concatenate([[0], np.cumsum(arr[:, 0])[:-1]]); tasks["t_off"] = np.concatenate([[0], np.cumsum(arr[:, 1])[:-1]])
    qa = np.concatenate(qparts) if qparts else np.zeros(0, np.uint8)
    ta = np.concatenate(tparts) if tparts else np.zeros(0, np.uint8)
    res, arena = aligner.align_batch(sc, qa, ta, tasks) if len(tasks) else (np.zeros(0, _abi.RESULT_DTYPE), np.zeros(0, np.uint32))
    if len(res) and (res["status"] != 0).any():
        raise ValueError("map_contigs: %d tasks were reset (scoring outside ksw2's int8 range)" % int((res["status"] != 0).sum()))

    def extension(tid):
        """(query bases, target bases, score, cigar) an extension task adds (ksw2_extz2_sse.c:292-301's end-point choice)"""
        if tid < 0:
            return 0, 0, 0, []
        r = res[tid]
        if int(r["reach_end"]):
            nq, nt, s_ = int(tasks["qlen"][tid]), int(r["mqe_t"]) + 1, int(r["mqe"])
        elif int(r["max_q"]) >= 0 and int(r["max_t"]) >= 0 and int(r["max"]) > 0:
            nq, nt, s_ = int(r["max_q"]) + 1, int(r["max_t"]) + 1, int(r["max"])
        else:
            return 0, 0, 0, []
        return nq, nt, s_, cigar_tuples(task_cigar(r, arena))

    out = []
    for i, ((chrom, wstart, _), (qname, _)) in enumerate(zip(windows, contigs)):
        recs = []
        qlen = len(qf[i])
        for ci, (c, pcs, qs, ids, left, right) in enumerate(plan[i]):
            core = cigar_tuples(stitch_cigars(pcs, np.array(ids, dtype=np.int32), res, arena)) if len(pcs) else []
            core_score = int(sum(int(res["score"][t]) for t in ids if t >= 0))
            for pc, t in zip(pcs, ids):           # a piece with an empty side is one gap
                if t < 0:
                    g = max(int(pc["q_end"] - pc["q_beg"]), int(pc["t_end"] - pc["t_beg"]))
                    core_score -= min(p.q + g * p.e, p.q2 + g * p.e2) if p.q2 >= 0 else p.q + g * p.e
            lq_, lt_, ls, lc = extension(left)
            rq_, rt_, rs, rc_ = extension(right)
            q0, q1 = int(c["q_beg"]) - lq_, int(c["q_end"]) + rq_
            t0 = int(c["t_beg"]) - lt_
            cigar = _merge_cigar([lc, core, rc_])
            supp = ci > 0
            clip = 5 if supp else 4
            full = ([(clip, q0)] if q0 else []) + cigar + ([(clip, qlen - q1)] if qlen - q1 else [])
            t1 = t0 + sum(n for op, n in cigar if op in (0, 2))
            nm = edit_distance_tag(cigar, qs[q0:q1], tt[i][t0:t1])
            seq_codes = qs[q0:q1] if supp else qs
            seq = "".join("ACGTN"[int(x)] for x in seq_codes)
            score = ls + core_score + rs
            # mapping quality (mm_set_mapq's form for an alignment with a DP score): identity x chain-size penalty x 40 x (1 - sub / score) x ln(score / a)
            blen = sum(n for op, n in cigar if op in (0, 1, 2))
            identity = max(0.0, (blen - nm) / blen) if blen else 0.0
            pen = min(1.0, 0.1 * int(c["n_anchors"]))
            sub = max(int(c["sub_score"]), 40)
            x = min(1.0, sub / max(int(c["score"]), 1))
            mapq = int(identity * pen * 40.0 * (1.0 - x) * np.log(max(score, p.a) / p.a)) if score > 0 else 0
            mapq = max(0, min(60, mapq))
            rec = SamRecord(qname, (16 if int(c["strand"]) else 0) | (2048 if supp else 0), chrom, int(wstart) + t0, mapq, full, seq,
                            {"NM": nm, "AS": score})
            recs.append((rec, q0, q1, nm))
        # SA tags (SAM spec: rname,pos,strand,CIGAR,mapQ,NM;) with minimap2's compact CIGAR (clip, M, one gap, clip)
        def sa_of(rec, q0, q1, nm):
            qspan, tspan = q1 - q0, rec.reference_end - rec.pos
            cg = ("%dS" % q0 if q0 else "") + "%dM" % min(qspan, tspan) + ("%dI" % (qspan - tspan) if qspan > tspan else "%dD" % (tspan - qspan) if tspan > qspan else "") + \
                 ("%dS" % (qlen - q1) if qlen - q1 else "")
            return "%s,%d,%s,%s,%d,%d;" % (rec.reference_name, rec.pos + 1, "-" if rec.is_reverse else "+", cg, rec.mapq, nm)
        if len(recs) > 1:
            for a_ in range(len(recs)):
                recs[a_][0].tags["SA"] = "".join(sa_of(*recs[b_]) for b_ in range(len(recs)) if b_ != a_)
        out.extend(r[0] for r in recs)
    return out


def records_from_results(windows, contigs, res, arena):
    """AlignedContig records of one global task per pair.  A task that ksw2 z-dropped (or whose band ran out) stops at its
    maximum cell: its CIGAR consumes only a prefix of the contig.  Such a record gets a trailing soft clip for the
    unaligned tail (so that the query-consuming operations add up to query_length, as every SAM record must), mapq 0 and
    zdropped = True: minimap2 would split or re-run that alignment (row f2), consumers must not take it for a full one."""
    out = []
    for i, ((chrom, start, _), (qname, qseq)) in enumerate(zip(windows, contigs)):
        if int(res[i]["status"]) != 0:
            raise ValueError("pair %d (%s): task status %d (scoring outside ksw2's int8 range)" % (i, qname, int(res[i]["status"])))
        cig = cigar_tuples(task_cigar(res[i], arena))
        ref_span = sum(n for op, n in cig if op in (0, 2))
        q_used = sum(n for op, n in cig if op in (0, 1, 4))
        dropped = bool(res[i]["zdropped"])
        mapq = 60
        if q_used < len(qseq):
            cig = cig + [(4, len(qseq) - q_used)]
            mapq = 0
        out.append(AlignedContig(qname, chrom, int(start), int(start) + ref_span, cig, False, mapq, len(qseq),
                                 int(res[i]["score"]), dropped))
    return out


def _merge_runs(sigs, rule):
    """Fold neighbouring signatures of one contig left to right with `rule(prev, cur) -> merged | None`."""
    if len(sigs) < 2:
        return sigs
    out = [sigs[0]]
    for cur in sigs[1:]:
        m = rule(out[-1], cur)
        if m is None:
            out.append(cur)
        else:
            out[-1] = m
    return out


def extract_sig_from_cigar(rec, min_svlen=30):
    """DEL / INS signatures of one aligned contig (extract_contig_signature_CCS.py:14-127).

    CIGAR walk: D >= min_svlen -> DEL at the reference offset; I >= min_svlen -> INS; soft clips advance
    the contig offset, a leading hard clip shifts contig coordinates.  Then, per contig, neighbouring INS
    are merged when both are long and close (>250 bp within 250 bp; >320 within 380; >100 within 250) and
    neighbouring DEL when both >150 bp start within 150 bp.  Returns (dels, inss, ref_end, contig_end)."""
    strand = "-" if rec.is_reverse else "+"
    hard = rec.cigar[0][1] if rec.cigar and rec.cigar[0][0] == 5 else 0
    ro, co = rec.pos, 0
    dels, inss = [], []
    for op, n in rec.cigar:
        if op == 0:
            ro += n; co += n
        elif op == 4:
            co += n
        elif op == 2:
            if n >= min_svlen:
                dels.append(Signature(rec.reference_name, "DEL", ro, n, rec.qname, co + hard, co + hard + 1, strand, "cigar", rec.mapq))
            ro += n
        elif op == 1:
            if n >= min_svlen:
                inss.append(Signature(rec.reference_name, "INS", ro, n, rec.qname, co + hard, co + hard + n, strand, "cigar", rec.mapq))
            co += n

    def ins_rule(a, b):
        near = abs(b.pos - a.pos)
        ok = (a.svlen > 250 and b.svlen > 250 and near < 250) or (a.svlen > 320 and b.svlen > 320 and near < 380) or \
             (a.svlen > 100 and b.svlen > 100 and near < 250)
        return a._replace(svlen=b.read_end - a.read_start, read_end=b.read_end) if ok else None

    def del_rule(a, b):
        ok = a.svlen > 150 and b.svlen > 150 and abs(b.pos - a.pos) < 150
        return a._replace(svlen=b.pos + b.svlen - a.pos, read_end=a.read_start + 1) if ok else None

    return _merge_runs(dels, del_rule), _merge_runs(inss, ins_rule), ro, co


def realign_regions_signatures(aligner, windows, contigs, preset="asm5", bw=2000, flag=0, min_svlen=30, zdrop=None):
    """The same alignment as realign_regions, but the CIGARs never leave the GPU: the DEL / INS signatures are
    extracted on the device (fsv_batch_signatures) and only they come back.  Returns a list of Signature in the
    order `signatures(realign_regions(...))` gives."""
    p = PRESETS[preset]
    sc = scoring_for(preset)
    q = [encode(s) for _, s in contigs]
    t = [encode(s) for _, _, s in windows]
    tasks = make_tasks([len(x) for x in q], [len(x) for x in t], ksw_band(bw), p.zdrop if zdrop is None else zdrop, 0, flag)
    qa = np.concatenate(q) if q else np.zeros(0, np.uint8)
    ta = np.concatenate(t) if t else np.zeros(0, np.uint8)
    b = aligner.batch(sc, qa, ta, tasks)
    try:
        b.run()
        sig = b.signatures(np.array([int(s) for _, s, _ in windows], dtype=np.int64), min_svlen)
    finally:
        b.close()
    out = []
    for r in sig:
        chrom = windows[int(r["task"])][0]
        qname = contigs[int(r["task"])][0]
        out.append(Signature(chrom, "INS" if int(r["svtype"]) else "DEL", int(r["pos"]), int(r["svlen"]), qname,
                             int(r["read_start"]), int(r["read_end"]), "+", "cigar", 60))
    return out


def signatures(records, min_svlen=30):
    """All DEL/INS signatures of a list of AlignedContig, in record order."""
    out = []
    for r in records:
        d, i, _, _ = extract_sig_from_cigar(r, min_svlen)
        out.extend(d); out.extend(i)
    return out
