"""Row f1: alignment records -> DEL/INS signatures -> per-haplotype clusters -> hp1/hp2 pairing -> VCF text.

Host-side mirror of what FocalSV runs on the BAM minimap2 produced, for the three platforms it distinguishes:
  focalsv/4_sv_calling/Dippav/extract_contig_signature_CCS.py  (HiFi;   extract_contig_sig_CCS, :673-760)
  focalsv/4_sv_calling/Dippav/extract_contig_signature_CLR.py  (CLR;    cigar-record filter :12-29,:384-386, split rule :327-343)
  focalsv/4_sv_calling/Dippav/extract_contig_signature_ONT.py  (ONT;    split rule :314-340)
Same inputs (records with pysam's fields, the chromosome sequence, the contig sequences, the VCF header lines), same
output (the text of dippav_variant_chr<N>.vcf, byte for byte: tests/test_sv_call.py compares with VCFs the reference's
own functions wrote, tests/golden/make_vcf_golden.py).  The stages:

  cigar signatures   extract_sig_from_cigar per record (:14-127)         -> hook.extract_sig_from_cigar (or, with the
                                                                            CIGARs still in HBM, fsv_batch_signatures)
  split signatures   extract_sig_from_split over consecutive records of one contig (:268-327)
  clustering         cluster_del / cluster_ins (:157-249): greedy leader clustering over position-sorted signatures
  merge              merge_all (:433-455)
  pairing            pair_sig (:504-559): hp1 x hp2 -> 1/1 or 0/1
  VCF                add_seq_to_sig + write_vcf (:598-670)

The reference's loops are quadratic in the number of signatures; here the position-sorted order its own sort_sig
establishes is used to look only at candidates within the distance thresholds (identical results: a candidate outside
them can never satisfy the rule), so a chromosome's worth of signatures clusters in milliseconds.
"""
from bisect import bisect_left

import numpy as np

from .hook import Signature, extract_sig_from_cigar

PLATFORMS = ("CCS", "CLR", "ONT")


def sort_by_pos(sigs):
    """sort_sig (:128-137): np.argsort of the positions (numpy's default sort, so ties land where the reference's do)."""
    if not sigs:
        return []
    return [sigs[i] for i in np.argsort([s.pos for s in sigs])]


def _is_sorted(sigs):
    return all(sigs[i].pos <= sigs[i + 1].pos for i in range(len(sigs) - 1))


def _similar_del(a, b, max_shift, min_overlap_ratio, min_size_similarity):
    """pair_del / the test inside cluster_del (:157-181, :476-497)."""
    s1, s2 = a.pos, b.pos
    e1, e2 = s1 + a.svlen, s2 + b.svlen
    overlap_ratio = (min(e1, e2) - max(s1, s2)) / min(e1 - s1, e2 - s2)
    size_similarity = min(a.svlen, b.svlen) / max(a.svlen, b.svlen)
    return abs(s1 - s2) <= max_shift and overlap_ratio >= min_overlap_ratio and size_similarity >= min_size_similarity


def _similar_ins(a, b, max_shift, min_size_similarity):
    """pair_ins / the test inside cluster_ins (:211-222, :466-474)."""
    return abs(a.pos - b.pos) <= max_shift and min(a.svlen, b.svlen) / max(a.svlen, b.svlen) >= min_size_similarity


def _leader_clusters(sigs, similar, max_shift):
    """The reference's clustering loop: walking the list in order, every signature that has no cluster yet becomes a
    leader and takes every later unclustered signature that is similar TO THE LEADER; a cluster is represented by its
    longest member (the first of equally long ones); clusters come out in the order of their leaders."""
    n = len(sigs)
    leader = [-1] * n
    windowed = _is_sorted(sigs)            # candidates further than max_shift to the right can be skipped only then
    out = []
    for i in range(n):
        if leader[i] != -1:
            continue
        leader[i] = i
        best = sigs[i]
        for j in range(i + 1, n):
            if windowed and sigs[j].pos - sigs[i].pos > max_shift:
                break
            if leader[j] == -1 and similar(sigs[i], sigs[j]):
                leader[j] = i
                if sigs[j].svlen > best.svlen:
                    best = sigs[j]
        out.append(best)
    return out


def cluster_del(sigs, max_shift=100, min_overlap_ratio=0.5, min_size_similarity=0.5):
    return _leader_clusters(sigs, lambda a, b: _similar_del(a, b, max_shift, min_overlap_ratio, min_size_similarity), max_shift)


def cluster_ins(sigs, max_shift=100, min_size_similarity=0.5):
    return _leader_clusters(sigs, lambda a, b: _similar_ins(a, b, max_shift, min_size_similarity), max_shift)


def _read_length(cigar):
    return sum(n for op, n in cigar if op in (0, 1, 4, 5))


def extract_sig_from_split(read1, read2, min_mapq, max_svlen, platform="CCS"):
    """Signatures from two consecutive alignments of ONE contig (:268-327; CLR :327-343; ONT :314-340): the first ends in
    a clip, the second starts with one, same strand; the reference gap and the contig gap between them give a DEL or an INS."""
    if read1.pos > read2.pos or read1.qname != read2.qname or read1.reference_name != read2.reference_name:
        raise AssertionError("extract_sig_from_split: records out of order or of different contigs")
    dels, inss = [], []
    c1, c2 = read1.cigar, read2.cigar
    if read1.is_reverse != read2.is_reverse or read1.mapq < min_mapq or read2.mapq < min_mapq or \
            c1[-1][0] not in (4, 5) or c2[0][0] not in (4, 5):
        return dels, inss
    rl1, rl2 = _read_length(c1), _read_length(c2)
    if rl1 != rl2:
        raise AssertionError("extract_sig_from_split: the two records of %s disagree on the contig length" % read1.qname)
    ref1e, ref2s = read1.reference_end, read2.pos
    rd1e, rd2s = rl1 - c1[-1][1], c2[0][1]
    diffdis = (ref2s - ref1e) - (rd2s - rd1e)
    diffolp = ref1e - ref2s
    strand = "-" if read1.is_reverse else "+"
    mq = "%d-%d" % (read1.mapq, read2.mapq)
    chrom, qn = read1.reference_name, read1.qname
    if abs(diffdis) > max_svlen:
        return dels, inss

    def ins_sig():
        svlen = abs(rd2s - rd1e + diffolp)
        pos_ref = int((ref1e + ref2s) / 2) if abs(diffolp) > 400 else ref2s
        return Signature(chrom, "INS", pos_ref, svlen, qn, rd1e - diffolp, rd2s, strand, "split-alignment", mq)

    if platform == "CCS":
        if diffolp < 30 and diffdis >= 30:
            dels.append(Signature(chrom, "DEL", ref1e, diffdis, qn, rd1e, rd2s, strand, "split-alignment", mq))
        elif diffolp < 3000 and diffdis >= 30:
            dels.append(Signature(chrom, "DEL", ref1e - diffdis, diffdis, qn, rd1e - diffdis, rd2s - diffdis, strand, "split-alignment", mq))
        elif diffolp < 3000 and diffdis <= -30:
            inss.append(ins_sig())
    else:
        r = 0.3 if platform == "CLR" else 0.5
        lo = r if platform == "CLR" else 0.8           # ONT tests Diffdis*0.8 <= Diffolp (:340), CLR Diffdis*r (:336)
        if diffdis >= 30:
            olp_read = rd1e - rd2s                      # (the DEL branch re-defines Diffolp on the contig, CLR :332)
            if -(diffdis * r) <= olp_read <= diffdis * r:
                dels.append(Signature(chrom, "DEL", ref1e, diffdis, qn, rd1e, rd2s, strand, "split-alignment", mq))
        elif diffdis * lo <= diffolp <= abs(diffdis) * r and diffdis <= -30:
            inss.append(ins_sig())
    return dels, inss


def _clr_record_ok(cigar):
    """CLR only (:12-29, :384-386): drop a record whose CIGAR is mostly insertions between short matches."""
    m = [n for op, n in cigar if op == 0]
    ins = sum(n for op, n in cigar if op == 1)
    ins_pct = ins / (sum(m) + ins)
    return ins_pct <= 0.13 or sum(m) / len(m) >= 200


def haplotype_signatures(reads, hp, min_cigar_mapq=50, min_split_mapq=50, platform="CCS", min_svlen=30):
    """extract_signature_one_hap (:457-471): `reads` = the chromosome's records in coordinate order (samfile.fetch)."""
    dels, inss = [], []
    for rd in reads:
        if hp in rd.qname and rd.mapq >= min_cigar_mapq and (platform != "CLR" or _clr_record_ok(rd.cigar)):
            d, i, ref_end, q_end = extract_sig_from_cigar(rd, min_svlen)
            if ref_end != rd.reference_end:
                raise AssertionError("record %s: the CIGAR ends at %d, reference_end says %d" % (rd.qname, ref_end, rd.reference_end))
            dels.extend(d); inss.extend(i)
    del_cigar = cluster_del(sort_by_pos(dels))
    ins_cigar = cluster_ins(sort_by_pos(inss))
    # contigs with more than one record (:378-398)
    count = {}
    for rd in reads:
        if hp in rd.qname and rd.mapq >= min_split_mapq:
            count[rd.qname] = count.get(rd.qname, 0) + 1
    multi = [name for name, c in count.items() if c > 1]
    by_name = {name: [] for name in multi}
    for rd in reads:
        if rd.qname in by_name and rd.mapq >= min_split_mapq:
            by_name[rd.qname].append(rd)
    sdel, sins = [], []
    for name in multi:
        rs = by_name[name]
        for a, b in zip(rs[:-1], rs[1:]):
            d, i = extract_sig_from_split(a, b, min_split_mapq, 50000, platform)
            sdel.extend(d); sins.extend(i)
    del_split = cluster_del(sort_by_pos(sdel))
    ins_split = cluster_ins(sort_by_pos(sins))
    ins_final = cluster_ins(sort_by_pos(ins_cigar + ins_split))
    del_final = cluster_del(sort_by_pos(del_cigar + del_split))
    return sort_by_pos(ins_final + del_final)


class PairedSignature(object):
    """One VCF row before the sequences are attached: the kept signature + GT and the per-haplotype INFO strings."""
    __slots__ = ("sig", "gt", "tig_region", "strands", "sources", "mapqs")

    def __init__(self, sig, gt, tig_region, strands, sources, mapqs):
        self.sig, self.gt, self.tig_region, self.strands, self.sources, self.mapqs = sig, gt, tig_region, strands, sources, mapqs

    @property
    def pos(self):
        return self.sig.pos

    def as_row(self):
        """the reference's list layout (indices 0..14)"""
        return list(self.sig) + [self.gt, self.tig_region, self.strands, self.sources, self.mapqs]


def pair_haplotypes(sig_hp1, sig_hp2, max_compare_dist=1000):
    """pair_sig (:504-559): every hp1 signature takes the first still unpaired hp2 signature of the same type that passes
    pair_del(200, 0.5, 0.5) / pair_ins(200, 0.5); a pair is homozygous (the longer allele is kept), the rest are 0/1."""
    n1, n2 = len(sig_hp1), len(sig_hp2)
    mate1, mate2 = [-1] * n1, [-1] * n2
    sorted2 = _is_sorted(sig_hp2)
    pos2 = [s.pos for s in sig_hp2]
    for i, a in enumerate(sig_hp1):
        j0 = bisect_left(pos2, a.pos - 200) if sorted2 else 0          # nothing left of pos - 200 can pass the shift test
        for j in range(j0, n2):
            b = sig_hp2[j]
            if b.pos - a.pos > max_compare_dist:
                break
            if a.chrom == b.chrom and a.svtype == b.svtype and mate2[j] == -1:
                ok = _similar_del(a, b, 200, 0.5, 0.5) if a.svtype == "DEL" else _similar_ins(a, b, 200, 0.5)
                if ok:
                    mate1[i], mate2[j] = j, i
                    break
    region = lambda s: "%s:%d-%d" % (s.qname, s.read_start, s.read_end)      # noqa: E731
    out = []
    for i, a in enumerate(sig_hp1):
        if mate1[i] == -1:
            out.append(PairedSignature(a, "0/1", region(a), a.strand, a.source, str(a.mapq)))
        else:
            b = sig_hp2[mate1[i]]
            keep = a if a.svlen > b.svlen else b
            out.append(PairedSignature(keep, "1/1", region(a) + "," + region(b), a.strand + "," + b.strand,
                                       a.source + "," + b.source, str(a.mapq) + "," + str(b.mapq)))
    for j, b in enumerate(sig_hp2):
        if mate2[j] == -1:
            out.append(PairedSignature(b, "0/1", region(b), b.strand, b.source, str(b.mapq)))
    return sort_by_pos(out)


_COMPLEMENT = {"N": "N", "A": "T", "T": "A", "G": "C", "C": "G"}


def reverse_complement(seq):
    return "".join(_COMPLEMENT[c] for c in seq.upper()[::-1])


def vcf_body(paired, chrom_seq, contigs):
    """add_seq_to_sig + the record loop of write_vcf (:598-670).  Rows whose contig is not in `contigs` are dropped, as
    the reference drops them.  Python's slice semantics are part of the format: an INS on the reverse strand takes
    contig[-read_end:-read_start] (empty when read_start is 0), POS - 1 indexes the chromosome string directly."""
    lines = []
    n_ins = n_del = 0
    for p in paired:
        s = p.sig
        if s.qname not in contigs:
            continue
        pos0 = s.pos - 1
        if s.svtype == "DEL":
            allele = chrom_seq[s.pos:s.pos + s.svlen]
            alt = chrom_seq[pos0]
            ref = alt + allele
            n_del += 1
            idx = n_del
        else:
            tig = contigs[s.qname]
            allele = reverse_complement(tig[-s.read_end:-s.read_start]) if s.strand == "-" else tig[s.read_start:s.read_end]
            ref = chrom_seq[pos0]
            alt = ref + allele
            n_ins += 1
            idx = n_ins
        info = "SVLEN=%d;SVTYPE=%s;TIG_REGION=%s;QUERY_STRAND=%s;SIG_SOURCE=%s;TIG_MAPQ=%s" % (
            len(alt) - len(ref), s.svtype, p.tig_region, p.strands, p.sources, p.mapqs)
        lines.append("%s\t%d\tdippav.%s.%s.%d\t%s\t%s\t%d\tPASS\t%s\tGT\t%s\n" % (
            s.chrom, s.pos, s.chrom, s.svtype, idx, ref.upper(), alt.upper(), 20, info, p.gt))
    return lines


def call_chromosome(reads, chrom_seq, contigs, header_lines, platform="CCS", min_cigar_mapq=50, min_split_mapq=50):
    """One chromosome of extract_contig_sig_{CCS,CLR,ONT} (:719-760): records in coordinate order -> VCF text.
    (The reference's __main__ passes min_cigar_mapq for both thresholds; both default to 50.)"""
    if platform not in PLATFORMS:
        raise ValueError("platform must be one of %s" % (PLATFORMS,))
    hp1 = haplotype_signatures(reads, "hp1", min_cigar_mapq, min_split_mapq, platform)
    hp2 = haplotype_signatures(reads, "hp2", min_cigar_mapq, min_split_mapq, platform)
    paired = pair_haplotypes(hp1, hp2)
    return "".join(list(header_lines) + vcf_body(paired, chrom_seq, contigs))


def call_variants(records, ref_seqs, contigs, header_lines, platform="CCS", chroms=None):
    """All chromosomes: {chrom: VCF text}.  `records` in any order; per chromosome they are put in coordinate order
    (stable), which is what `samtools sort | pysam fetch` hands the reference."""
    by_chrom = {}
    for r in records:
        by_chrom.setdefault(r.reference_name, []).append(r)
    out = {}
    for chrom in (chroms if chroms is not None else sorted(by_chrom)):
        reads = sorted(by_chrom.get(chrom, []), key=lambda r: r.pos)
        out[chrom] = call_chromosome(reads, ref_seqs[chrom], contigs, header_lines, platform)
    return out
