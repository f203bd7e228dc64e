"""Row f1 to the VCF: focalsv_b200.sv_call against VCF text written by the REFERENCE'S OWN code.

tests/golden/vcf_golden.json.gz was produced by tests/golden/make_vcf_golden.py: the reference's unmodified
extract_contig_sig_{CCS,CLR,ONT} (focalsv/4_sv_calling/Dippav/extract_contig_signature_*.py) ran on SAM files written by
focalsv_b200.dropin, with dropin.AlignmentFile standing in for pysam.  Here the same inputs (regenerated from seeds) go
through sv_call and must give the same VCF byte for byte: CPU tests with the oracle's alignments / the frozen records, GPU
tests with the GPU's alignments (which closes the chain  GPU CIGARs -> signatures -> clusters -> pairs -> VCF).
"""
import gzip
import json
import os

import numpy as np
import pytest

import vcf_cases
from focalsv_b200 import dropin, hook, sv_call
from util import OracleRunner

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    with gzip.open(os.path.join(HERE, "golden", "vcf_golden.json.gz"), "rt") as fh:
        return json.load(fh)


def _case(golden, kind, platform, seed):
    return [c for c in golden["cases"] if (c["kind"], c["platform"], c["seed"]) == (kind, platform, seed)][0]


def _frozen(c):
    return [dropin.SamRecord(q, fl, rn, pos, mq, cig, None) for q, fl, rn, pos, mq, cig in c["records"]]


def _vcf(golden, platform, chrom, tigs, recs):
    reads = sorted(recs, key=lambda r: r.pos)
    return sv_call.call_chromosome(reads, vcf_cases.ascii_of(chrom), tigs, golden["header"].splitlines(True), platform)


@pytest.mark.parametrize("platform,seed", [(p, s) for k, p, s in vcf_cases.CASES if k == "records"])
def test_records_to_vcf_matches_the_reference(golden, platform, seed):
    """Random records (dense DEL/INS, contigs split into clipped pieces, all mapq / strand filters) -> VCF text."""
    c = _case(golden, "records", platform, seed)
    chrom, tigs, recs = vcf_cases.record_case(seed)
    assert [[r.qname, r.flag, r.reference_name, r.pos, r.mapq, [list(x) for x in r.cigar]] for r in recs] == c["records"]
    assert _vcf(golden, platform, chrom, tigs, recs) == c["vcf"]


@pytest.mark.parametrize("platform,seed", [(p, s) for k, p, s in vcf_cases.CASES if k == "regions"])
def test_regions_to_vcf_matches_the_reference_cpu(golden, oracle, platform, seed):
    """Scaled config: contigs of 8 regions x 2 haplotypes aligned by the oracle arm -> records -> VCF text."""
    c = _case(golden, "regions", platform, seed)
    chrom, windows, contigs = vcf_cases.region_case(seed, platform)
    aligned = hook.realign_regions(OracleRunner(oracle), windows, contigs, preset="asm5", bw=2000)
    recs = dropin.sam_records(aligned, windows, contigs)
    assert [[r.qname, r.flag, r.reference_name, r.pos, r.mapq, [list(x) for x in r.cigar]] for r in recs] == c["records"]
    tigs = {n: vcf_cases.ascii_of(s) for n, s in contigs}
    vcf = _vcf(golden, platform, chrom, tigs, recs)
    assert vcf == c["vcf"]
    assert vcf.count("\n") > len(golden["header"].splitlines()) + 10 and "\t1/1\n" in vcf and "\t0/1\n" in vcf


def test_sam_round_trip_and_pysam_shape(tmp_path, golden):
    """write_sam -> AlignmentFile: the records the reference's consumers would fetch carry the same fields."""
    c = _case(golden, "records", "CCS", 21)
    recs = _frozen(c)
    for r in recs[:5]:
        r.tags.update({"NM": 7, "AS": 1234}); r.seq = "ACGT"
    path = str(tmp_path / "x.sam")
    dropin.write_sam(path, recs, [(vcf_cases.CHROM, 300000)])
    back = list(dropin.AlignmentFile(path).fetch(vcf_cases.CHROM))
    assert len(back) == len(recs) and all(a.pos <= b.pos for a, b in zip(back[:-1], back[1:]))
    key = lambda r: (r.pos, r.qname, r.flag)      # noqa: E731
    for a, b in zip(sorted(back, key=key), sorted(recs, key=key)):
        assert (a.qname, a.flag, a.reference_name, a.pos, a.mapq, a.cigar, a.reference_end, a.is_reverse, a.seq, a.tags) == \
               (b.qname, b.flag, b.reference_name, b.pos, b.mapq, b.cigar, b.reference_end, b.is_reverse, b.seq, b.tags)
        assert a.query_name == a.qname and a.reference_start == a.pos and a.mapping_quality == a.mapq and a.cigartuples == a.cigar
    assert back[0].cigarstring.endswith(tuple("MIDSH")) and list(dropin.AlignmentFile(path).fetch("chrX")) == []


def test_clustering_windowed_equals_quadratic():
    """The position window used for speed cannot change a cluster: compare with the reference's all-pairs loop restated naively."""
    rng = np.random.default_rng(3)

    def naive(sigs, similar):
        cluster = [-1] * len(sigs)
        for i in range(len(sigs)):
            if cluster[i] == -1:
                cluster[i] = i
                for j in range(len(sigs)):
                    if cluster[j] == -1 and similar(sigs[i], sigs[j]):
                        cluster[j] = i
        out = []
        for lead in dict.fromkeys(cluster):
            members = [s for s, cidx in zip(sigs, cluster) if cidx == lead]
            best = members[0]
            for m in members:
                if m.svlen > best.svlen:
                    best = m
            out.append(best)
        return out

    for _ in range(30):
        n = int(rng.integers(0, 120))
        sigs = [hook.Signature("chr1", "DEL", int(rng.integers(0, 3000)), int(rng.integers(30, 400)), "c%d" % k, 0, 1, "+", "cigar", 60) for k in range(n)]
        srt = sv_call.sort_by_pos(sigs)
        assert sv_call.cluster_del(srt) == naive(srt, lambda a, b: sv_call._similar_del(a, b, 100, 0.5, 0.5))
        assert sv_call.cluster_ins(srt) == naive(srt, lambda a, b: sv_call._similar_ins(a, b, 100, 0.5))
        assert sv_call.cluster_del(sigs) == naive(sigs, lambda a, b: sv_call._similar_del(a, b, 100, 0.5, 0.5))      # unsorted input: no window


@pytest.mark.gpu
@pytest.mark.parametrize("platform,seed", [(p, s) for k, p, s in vcf_cases.CASES if k == "regions"])
def test_regions_to_vcf_matches_the_reference_gpu(golden, aligner, tmp_path, platform, seed):
    """The same scaled configs with the alignments made on the GPU, through the FASTA-in / SAM-out drop-in: identical VCF."""
    c = _case(golden, "regions", platform, seed)
    chrom, windows, contigs = vcf_cases.region_case(seed, platform)
    ref_fa = str(tmp_path / "ref_chr21.fa")
    with open(ref_fa, "w") as fh:
        s = vcf_cases.ascii_of(chrom)
        fh.write(">%s\n" % vcf_cases.CHROM + "\n".join(s[i:i + 80] for i in range(0, len(s), 80)) + "\n")
    hp_fa, wins = [], {}
    for h in (1, 2):
        p = str(tmp_path / ("hp%d.fa" % h)); hp_fa.append(p)
        with open(p, "w") as fh:
            for (name, q), (_, start, t) in zip(contigs, windows):
                if "_hp%d_" % h in name:                     # original assembler names; align_fastas renames them contig_hp<h>_<n>
                    fh.write(">ptg%s %s:%d-%d\n%s\n" % (name.split("_")[-1], vcf_cases.CHROM, start, start + len(t), vcf_cases.ascii_of(q)))
    sam = str(tmp_path / "assemblies.sorted.sam")
    recs = dropin.align_fastas(aligner, ref_fa, hp_fa, None, sam, preset="asm5", bw=2000)
    got = sorted([[r.qname, r.flag, r.reference_name, r.pos, r.mapq, [list(x) for x in r.cigar]] for r in recs])
    assert got == sorted(c["records"])
    fetched = list(dropin.AlignmentFile(sam).fetch(vcf_cases.CHROM))
    tigs = {n: vcf_cases.ascii_of(s) for n, s in contigs}
    vcf = sv_call.call_chromosome(fetched, vcf_cases.ascii_of(chrom), tigs, golden["header"].splitlines(True), platform)
    assert vcf == c["vcf"]
    nm = {r.qname: r.get_tag("NM") for r in fetched}
    assert all(v >= 0 for v in nm.values()) and all(r.seq and len(r.seq) == r.query_length for r in fetched)
