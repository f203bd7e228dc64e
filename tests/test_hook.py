"""Pipeline hook: CIGAR -> DEL/INS signatures (extract_contig_signature_CCS.py:14-127 semantics), and the
end-to-end region call on the GPU giving the same records as the oracle-backed path."""
import numpy as np
import pytest

from focalsv_b200 import _abi, hook, synth
from focalsv_b200.presets import PRESETS, ksw_band, scoring_for


def _regions(seed, n=4, L=9000):
    rng = np.random.default_rng(seed)
    windows, contigs, truth = [], [], []
    for i in range(n):
        ref = synth.random_seq(rng, L + 500 * i)
        q, svs = synth.plant_svs(rng, ref, 3, max_net=1200, max_len=900)
        q = synth.mutate(rng, q, 0.0006, 0.0002, 0.0002)
        windows.append(("chr21", 1000000 + 20000 * i, ref)); contigs.append(("contig_hp1_%d" % i, q)); truth.append(svs)
    return windows, contigs, truth


def _oracle_records(O, windows, contigs, preset="asm5"):
    p = PRESETS[preset]; sc = scoring_for(preset)
    recs = []
    for (chrom, start, t), (qn, q) in zip(windows, contigs):
        r, cig = O.extd2(q, t, sc, w=ksw_band(2000), zdrop=p.zdrop, flag=0)
        cg = hook.cigar_tuples(cig)
        ref_end = start + sum(n for op, n in cg if op in (0, 2))
        q_used = sum(n for op, n in cg if op in (0, 1))
        mapq = 60
        if q_used < len(q):         # z-dropped: soft clip for the unaligned tail, mapq 0 (hook.records_from_results)
            cg = cg + [(4, len(q) - q_used)]; mapq = 0
        recs.append(hook.AlignedContig(qn, chrom, start, ref_end, cg, False, mapq, len(q), int(r["score"]), bool(r["zdropped"])))
    return recs


def test_cigar_walk_rules():
    rec = hook.AlignedContig("c", "chr1", 100, 0, [(5, 7), (0, 50), (2, 40), (0, 10), (1, 29), (0, 5), (1, 300), (0, 20), (1, 280), (0, 3), (2, 200), (0, 100), (2, 160)],
                             False, 60, 0, 0, False)
    dels, inss, ro, co = hook.extract_sig_from_cigar(rec, 30)
    assert [(d.pos, d.svlen) for d in dels] == [(150, 40), (228, 200), (528, 160)]       # starts 300 bp apart: not merged
    rec2 = rec._replace(cigar=[(0, 50), (2, 200), (0, 100), (2, 160), (0, 9)])
    d2, _, _, _ = hook.extract_sig_from_cigar(rec2, 30)
    assert [(d.pos, d.svlen) for d in d2] == [(150, 200), (450, 160)]     # the reference compares START positions, so two >150 bp DELs never merge
    assert len(inss) == 1 and inss[0].svlen == (50 + 10 + 29 + 5 + 300 + 20 + 280) - (50 + 10 + 29 + 5)   # two long INS 20 bp apart merge
    assert inss[0].read_start == 50 + 10 + 29 + 5 + 7                                      # hard clip shifts contig coordinates


def test_planted_svs_are_recovered_from_oracle_alignments(oracle):
    windows, contigs, truth = _regions(5)
    recs = _oracle_records(oracle, windows, contigs)
    assert sum(1 for r in recs if not r.zdropped) >= 2
    for rec, svs, (chrom, start, _) in zip(recs, truth, windows):
        if rec.zdropped:            # ksw2's z-drop ends an alignment inside a noisy planted SV now and then
            continue
        sigs = hook.signatures([rec])
        want = sorted((t, L) for _, t, L in svs if L >= 30)
        got = sorted((s.svtype, s.svlen) for s in sigs)
        assert len(got) == len(want)
        for (gt, gl), (wt, wl) in zip(sorted(got, key=lambda x: x[1]), sorted(want, key=lambda x: x[1])):
            assert gt == wt and abs(gl - wl) <= 12


@pytest.mark.gpu
def test_gpu_region_call_gives_identical_records_and_signatures(oracle, aligner):
    windows, contigs, _ = _regions(6, n=6)
    want = _oracle_records(oracle, windows, contigs)
    got = hook.realign_regions(aligner, windows, contigs, preset="asm5", bw=2000)
    assert got == want
    assert hook.signatures(got) == hook.signatures(want)


@pytest.mark.gpu
def test_level1_entry_point_gives_the_same_records(oracle, aligner):
    """fsv_realign_regions (SURVEY 8b level 1): windows addressed inside ONE reference arena, overlapping allowed."""
    rng = np.random.default_rng(9)
    ref = synth.random_seq(rng, 60000)
    regions, contigs, windows = [], [], []
    for i, (s, e) in enumerate([(1000, 11000), (8000, 20500), (30000, 39000), (30000, 39000), (59000, 60000)]):
        q, _ = synth.plant_svs(rng, ref[s:e], 2, max_net=900, max_len=700)
        q = synth.mutate(rng, q, 0.0006, 0.0002, 0.0002)
        regions.append(("chr21", s, e)); contigs.append(("contig_hp%d_%d" % (1 + i % 2, i), q)); windows.append(("chr21", s, ref[s:e]))
    want = _oracle_records(oracle, windows, contigs)
    got = hook.realign_regions_abi(aligner, ref, regions, contigs, preset="asm5", bw=2000)
    assert got == want
    assert hook.realign_regions_abi(aligner, ref, [], [], preset="asm5") == []
    from focalsv_b200.api import FsvError
    with pytest.raises(FsvError):
        hook.realign_regions_abi(aligner, ref, [("chr21", 100, 70000)], contigs[:1])          # window past the reference
    with pytest.raises(FsvError):
        hook.realign_regions_abi(aligner, ref, regions[:1], contigs[:1], preset="no-such-preset")


class _OracleRunner(object):
    """align_batch on the CPU oracle: lets the host side of row f2 (seeding, chaining, stitching) be tested without a GPU."""
    def __init__(self, O):
        self.O = O

    def align_batch(self, sc, qa, ta, tasks):
        return self.O.run_batch(sc, qa, ta, tasks, threads=8)


def _consumed(cigar):
    return sum(n for op, n in cigar if op in (0, 1)), sum(n for op, n in cigar if op in (0, 2))


def test_chain_pieces_tile_the_pair_and_isolate_the_svs():
    from focalsv_b200 import api
    rng = np.random.default_rng(77)
    ref = synth.random_seq(rng, 120000)
    q, svs = synth.plant_svs(rng, ref, 8, max_net=9000, max_len=9000)
    q = synth.mutate(rng, q, 0.001, 0.0002, 0.0002)
    pcs, score, n_anchor = api.chain_pieces(q, ref, 19, 19, 50, 100000, 200)
    assert n_anchor > 5000 and score > 50000
    assert pcs["q_beg"][0] == 0 and pcs["t_beg"][0] == 0 and pcs["q_end"][-1] == len(q) and pcs["t_end"][-1] == len(ref)
    assert (pcs["q_beg"][1:] == pcs["q_end"][:-1]).all() and (pcs["t_beg"][1:] == pcs["t_end"][:-1]).all()
    dq = (pcs["q_end"] - pcs["q_beg"]).astype(np.int64); dt = (pcs["t_end"] - pcs["t_beg"]).astype(np.int64)
    assert int((dq * dt).sum()) < 0.05 * api.task_cells(len(q), len(ref), 3001)          # a small fraction of the band-3001 DP of the whole pair
    nets = sorted(int(x) for x in (dq - dt) if abs(int(x)) >= 50)
    want = sorted((L if t == "INS" else -L) for _, t, L in svs if L >= 50)
    assert len(nets) == len(want) and all(abs(a - b) <= 12 for a, b in zip(nets, want)), (nets, want)
    # degenerate inputs: nothing shared -> one piece; empty sides -> a pure gap piece
    a = synth.random_seq(rng, 500); b = synth.random_seq(rng, 700)
    pcs, _, n_anchor = api.chain_pieces(a, b)
    assert n_anchor == 0 and len(pcs) == 1 and tuple(pcs[0]) == (0, 500, 0, 700)
    pcs, _, _ = api.chain_pieces(a, np.zeros(0, np.uint8))
    assert len(pcs) == 1 and tuple(pcs[0]) == (0, 500, 0, 0)



def _naive_minimizers(s, k, w):
    """(w, k)-minimizers of the forward strand exactly as fsv_chain.cu defines them (hash mix64 of the 2-bit k-mer, the smallest of every
    window of w consecutive k-mers, leftmost on ties, windows never span a wildcard), the slow obvious way: end positions of the chosen k-mers."""
    M = (1 << 64) - 1

    def mix64(x):
        x ^= x >> 33; x = (x * 0xff51afd7ed558ccd) & M; x ^= x >> 33; x = (x * 0xc4ceb9fe1a85ec53) & M; x ^= x >> 33
        return x
    out, run = set(), []            # run: (hash, end position) of the k-mers since the last wildcard
    valid, kmer, mask = 0, 0, (1 << (2 * k)) - 1
    for i, c in enumerate(int(x) for x in s):
        if c > 3:
            valid, run = 0, []
            continue
        kmer = ((kmer << 2) | c) & mask
        valid += 1
        if valid < k:
            continue
        run.append((mix64(kmer), i))
        if len(run) >= w:
            out.add(min(run[-w:])[1])
    return out


def test_piece_boundaries_are_shared_minimizers():
    """Every cut between two pieces is the END of a k-mer that both sequences share and that is a minimizer of both (what the sketch + anchor
    join must deliver whatever their implementation: the O(n) monotonic-queue sketch is checked against the obvious O(n w) definition)."""
    from focalsv_b200 import api
    rng = np.random.default_rng(5)
    for (k, w), L in (((19, 19), 6000), ((15, 10), 4000), ((19, 10), 3000)):
        ref = synth.random_seq(rng, L)
        q, _ = synth.plant_svs(rng, ref, 3, max_net=600, max_len=500)
        q = synth.mutate(rng, q, 0.01, 0.003, 0.003)
        q = q.copy(); q[rng.integers(0, len(q), 4)] = 4                    # a few wildcards: windows restart behind them
        ref = ref.copy(); ref[L // 2:L // 2 + 120] = 0                      # a homopolymer stretch: ties, leftmost wins
        pcs, _, n_anchor = api.chain_pieces(q, ref, k, w, 50, 100000, 200)
        assert n_anchor > 20 and len(pcs) > 5
        mq, mt = _naive_minimizers(q, k, w), _naive_minimizers(ref, k, w)
        for pc in pcs[:-1]:                                                # the last piece ends at the sequence ends
            qe, te = int(pc["q_end"]), int(pc["t_end"])
            if qe < k or te < k:                                           # (the leading piece before the first anchor starts at its k-mer's first base)
                continue
            same = np.array_equal(q[qe - k:qe], ref[te - k:te])
            if not same:                                                   # the first cut is the START of the first anchor: its k-mer lies behind it
                assert np.array_equal(q[qe:qe + k], ref[te:te + k]) and (qe + k - 1) in mq and (te + k - 1) in mt, (k, w, qe, te)
                continue
            assert (qe - 1) in mq and (te - 1) in mt, (k, w, qe, te)

def test_chained_alignment_is_a_valid_global_alignment_and_recovers_the_svs(oracle):
    windows, contigs, truth = _regions(21, n=3, L=20000)
    recs = hook.realign_regions_chained(_OracleRunner(oracle), windows, contigs, preset="asm5", bw=2000)
    for rec, (chrom, start, t), (qn, q), svs in zip(recs, windows, contigs, truth):
        assert _consumed(rec.cigar) == (len(q), len(t)) and rec.pos == start and rec.reference_end == start + len(t)
        got = sorted((s.svtype, s.svlen) for s in hook.signatures([rec]))
        want = sorted((ty, L) for _, ty, L in svs if L >= 30)
        assert len(got) == len(want)
        for (gt, gl), (wt, wl) in zip(sorted(got, key=lambda x: x[1]), sorted(want, key=lambda x: x[1])):
            assert gt == wt and abs(gl - wl) <= 12
    # re-scoring the stitched CIGAR with the preset gives the reported score
    p = PRESETS["asm5"]
    for rec, (_, _, t), (_, q) in zip(recs, windows, contigs):
        qi = ti = sc = 0
        qq, tt = hook.encode(q), hook.encode(t)
        for op, n in rec.cigar:
            if op == 0:
                eq = qq[qi:qi + n] == tt[ti:ti + n]
                sc += int(eq.sum()) * p.a - int((~eq).sum()) * p.b; qi += n; ti += n
            elif op == 1:
                sc -= min(p.q + n * p.e, p.q2 + n * p.e2); qi += n
            else:
                sc -= min(p.q + n * p.e, p.q2 + n * p.e2); ti += n
        assert sc == rec.score


@pytest.mark.gpu
def test_gpu_chained_alignment_equals_the_oracle_run(oracle, aligner):
    windows, contigs, _ = _regions(22, n=4, L=30000)
    want = hook.realign_regions_chained(_OracleRunner(oracle), windows, contigs, preset="asm5", bw=2000)
    got = hook.realign_regions_chained(aligner, windows, contigs, preset="asm5", bw=2000)
    assert got == want


def test_library_presets_equal_the_python_table():
    """fsv_preset_lookup is host-only: the library's preset table and ksw_gen_simple_mat restatement against presets.py."""
    from focalsv_b200 import api
    for name, p in PRESETS.items():
        f, sc = api.preset_lookup(name)
        assert f == {k: getattr(p, k) for k in f}
        want = scoring_for(name)
        assert (sc.m, sc.q, sc.e, sc.q2, sc.e2) == (want.m, want.q, want.e, want.q2, want.e2)
        assert list(sc.mat)[:25] == list(want.mat)[:25]
    with pytest.raises(api.FsvError):
        api.preset_lookup("asm20")


def test_signature_rules_match_the_references_own_function():
    """tests/golden/sig_golden.json holds inputs and outputs of the reference's own extract_sig_from_cigar
    (extract_contig_signature_CCS.py:14-127, run by tests/golden/make_sig_golden.py): clips, adjacent I/D,
    every merge rule."""
    import json, os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sig_golden.json")))
    assert len(g["cases"]) >= 100
    merged = 0
    for c in g["cases"]:
        r = c["record"]
        rec = hook.AlignedContig(r["qname"], r["reference_name"], r["pos"], 0, [tuple(x) for x in r["cigar"]],
                                 r["is_reverse"], r["mapq"], 0, 0, False)
        d, i, ro, co = hook.extract_sig_from_cigar(rec, c["min_svlen"])
        assert [list(x) for x in d] == c["dels"] and [list(x) for x in i] == c["inss"]
        assert (ro, co) == (c["ref_end"], c["contig_end"])
        n_raw = sum(1 for op, n in rec.cigar if op in (1, 2) and n >= c["min_svlen"])
        merged += n_raw - len(d) - len(i)
    assert merged > 20          # the merge rules were exercised


def _close_sv_regions(seed, n=5):
    """Contigs whose planted events sit close together, so that the per-contig merge rules fire."""
    rng = np.random.default_rng(seed)
    windows, contigs = [], []
    for i in range(n):
        ref = synth.random_seq(rng, 9000 + 700 * i)
        k = 3000
        if i % 2 == 0:      # two long insertions 60 bp apart, then a deletion
            q = np.concatenate([ref[:k], synth.random_seq(rng, 400), ref[k:k + 60], synth.random_seq(rng, 350), ref[k + 60:6000], ref[6300:]])
        else:               # two long deletions whose starts are 100 bp apart
            q = np.concatenate([ref[:k], ref[k + 200:k + 300], ref[k + 500:]])
        q = synth.mutate(rng, q, 0.0005, 0.0002, 0.0002)
        windows.append(("chr%d" % (1 + i), 5000000 + 31337 * i, ref)); contigs.append(("contig_hp2_%d" % i, q))
    return windows, contigs


@pytest.mark.gpu
def test_device_signatures_equal_the_host_rules(aligner):
    """fsv_batch_signatures (CIGARs stay in HBM) against extract_sig_from_cigar on the fetched CIGARs."""
    # z-drop 2000: with asm5's 200 two large events this close end the alignment (minimap2 would re-chain there)
    for windows, contigs in (_regions(8, n=6)[:2], _close_sv_regions(9)):
        recs = hook.realign_regions(aligner, windows, contigs, preset="asm5", bw=2000, zdrop=2000)
        for min_svlen in (30, 50):
            want = hook.signatures(recs, min_svlen)
            got = hook.realign_regions_signatures(aligner, windows, contigs, preset="asm5", bw=2000, min_svlen=min_svlen, zdrop=2000)
            assert got == want
            assert len(want) >= len(windows)
    # the second set really merged something: fewer signatures than qualifying CIGAR operations
    raw = sum(1 for r in recs for op, n in r.cigar if op in (1, 2) and n >= 30)
    assert len(hook.signatures(recs, 30)) < raw


@pytest.mark.gpu
def test_device_signatures_empty_and_capacity(aligner):
    from focalsv_b200.presets import scoring_for as sf
    from focalsv_b200.api import make_tasks
    seq = synth.random_seq(np.random.default_rng(1), 500)
    b = aligner.batch(sf("asm5"), seq, seq, make_tasks([500], [500], 100, 200))
    try:
        with pytest.raises(Exception):
            b.signatures()                      # not run yet: FSV_ERR_STATE
        b.run()
        assert len(b.signatures(np.array([7], dtype=np.int64))) == 0      # identical sequences: no signature
    finally:
        b.close()


def test_multi_aligner_gathers_in_task_order(oracle):
    """In-process multi-device path: LPT bins over the devices, host gather in the caller's order.  Two oracle-backed
    'devices' must give exactly what one gives (host logic only; on a GPU box the same class wraps api.Aligner(dev))."""
    from util import OracleRunner
    windows, contigs, _ = _regions(21, n=9, L=3000)
    one = hook.realign_regions(OracleRunner(oracle), windows, contigs, preset="asm5", bw=500)
    multi = hook.MultiAligner([OracleRunner(oracle, threads=2), OracleRunner(oracle, threads=2), OracleRunner(oracle, threads=2)])
    three = hook.realign_regions(multi, windows, contigs, preset="asm5", bw=500)
    assert one == three
    assert hook.realign_regions(hook.MultiAligner([OracleRunner(oracle)]), windows[:2], contigs[:2], preset="asm5", bw=500) == one[:2]


@pytest.mark.gpu
def test_multi_aligner_on_the_visible_gpus(oracle):
    from focalsv_b200 import api
    n_dev = max(1, min(api.load_library().fsv_device_count(), 4))
    windows, contigs, _ = _regions(22, n=7, L=6000)
    want = _oracle_records(oracle, windows, contigs)
    multi = hook.MultiAligner.on_devices(list(range(n_dev)) if n_dev > 1 else [0, 0])      # one GPU: two contexts on it
    try:
        assert hook.realign_regions(multi, windows, contigs, preset="asm5", bw=2000) == want
    finally:
        multi.close()


def _mm2_cases(seed=31):
    rng = np.random.default_rng(seed)
    ref = synth.random_seq(rng, 70000)
    q, svs = synth.plant_svs(rng, ref[5000:60000], 5, max_net=6000, max_len=5000)
    q = synth.mutate(rng, q, 0.001, 0.0002, 0.0002)
    inv = np.concatenate([q[:20000], (3 - q[20000:29000][::-1]).astype(np.uint8), q[29000:]])
    windows = [("chr7", 2000000, ref)] * 3
    contigs = [("ctg_fwd", q), ("ctg_rev", (3 - q[::-1]).astype(np.uint8)), ("ctg_inv", inv)]
    return ref, q, svs, windows, contigs


def _check_mm2_records(recs, ref, q, svs, windows, contigs):
    from focalsv_b200 import dropin
    by = {}
    for r in recs:
        by.setdefault(r.qname, []).append(r)
    fwd, rev, inv = by["ctg_fwd"], by["ctg_rev"], by["ctg_inv"]
    assert len(fwd) == 1 and len(rev) == 1 and fwd[0].flag == 0 and rev[0].flag == 16
    # the reverse-complemented contig aligns to the same place with the same CIGAR; SEQ is given on the reference strand
    assert (fwd[0].pos, fwd[0].cigar, fwd[0].tags["NM"]) == (rev[0].pos, rev[0].cigar, rev[0].tags["NM"]) and fwd[0].seq == rev[0].seq
    assert fwd[0].mapq == 60 and abs(fwd[0].pos - (2000000 + 5000)) < 100 and abs(fwd[0].reference_end - (2000000 + 60000)) < 100
    # every record is a valid SAM record: query-consuming operations add up to the contig, SEQ matches, NM re-computes
    for r in recs:
        qlen = len(dict(contigs)[r.qname])
        assert sum(n for op, n in r.cigar if op in (0, 1, 4, 5)) == qlen
        assert len(r.seq) == sum(n for op, n in r.cigar if op in (0, 1, 4))
        assert r.cigar[0][0] != 2 and r.cigar[-1][0] != 2 and all(n > 0 for _, n in r.cigar)
    # the planted SVs come out of the primary's CIGAR (extension ends: no spurious gap at the window's corners)
    sigs = hook.signatures([hook.AlignedContig(fwd[0].qname, fwd[0].reference_name, fwd[0].pos, fwd[0].reference_end, fwd[0].cigar, False, 60, 0, 0, False)])
    got = sorted((s.svtype, s.svlen) for s in sigs)
    want = sorted((t, L) for _, t, L in svs if L >= 30)
    assert len(got) == len(want) and all(a[0] == b[0] and abs(a[1] - b[1]) <= 12 for a, b in zip(sorted(got, key=lambda x: x[1]), sorted(want, key=lambda x: x[1])))
    # the inversion: a primary, two supplementary records, the inverted piece on the other strand, SA tags naming each other
    assert len(inv) == 3 and sorted(r.flag for r in inv) == [0, 2048, 2064]
    mid = [r for r in inv if r.flag == 2064][0]
    assert abs((mid.reference_end - mid.pos) - 9000) < 800 and all(r.has_tag("SA") and r.get_tag("SA").count(";") == 2 for r in inv)
    spans = sorted((r.pos, r.reference_end) for r in inv)
    assert all(spans[i][1] <= spans[i + 1][0] + 50 for i in range(2))           # the three pieces tile the locus


def test_map_contigs_strand_split_extension_cpu(oracle):
    """Row f2 second version on the oracle arm (host logic: chaining on both strands, hole splitting, extension ends, tags)."""
    from util import OracleRunner
    ref, q, svs, windows, contigs = _mm2_cases()
    _check_mm2_records(hook.map_contigs(OracleRunner(oracle), windows, contigs, preset="asm5", bw=2000), ref, q, svs, windows, contigs)


@pytest.mark.gpu
def test_map_contigs_gpu_equals_oracle_arm(oracle, aligner):
    from util import OracleRunner
    ref, q, svs, windows, contigs = _mm2_cases(32)
    want = hook.map_contigs(OracleRunner(oracle), windows, contigs, preset="asm5", bw=2000)
    got = hook.map_contigs(aligner, windows, contigs, preset="asm5", bw=2000)
    assert [r.to_sam() for r in got] == [r.to_sam() for r in want]
    _check_mm2_records(got, ref, q, svs, windows, contigs)
