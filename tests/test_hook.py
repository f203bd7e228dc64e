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
        recs.append(hook.AlignedContig(qn, chrom, start, start + sum(n for op, n in cg if op in (0, 2)), cg, False, 60,
                                       len(q), int(r["score"]), bool(r["zdropped"])))
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
