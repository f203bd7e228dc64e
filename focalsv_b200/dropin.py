"""The call site as an executable drop-in: FASTA in, SAM out, pysam-shaped records for the reference's consumers.

What FocalSV does today (focalsv/4_sv_calling/Dippav/DipPAV_variant_call.py:83-112):

    reformat_fasta(hp1) ; reformat_fasta(hp2)            # headers become contig_hp1_<n> / contig_hp2_<n>
    cat hp1.fa hp2.fa > assemblies.fa
    minimap2 -a -x asm5 --cs -r2k -t N ref_chrN.fa assemblies.fa | samtools sort > assemblies.sorted.bam
    samtools index assemblies.sorted.bam
    ... pysam.AlignmentFile(bam).fetch(chr_name) ...     # extract_contig_signature_CCS.py:343-356, 378-398, 726-736

and the same pattern in call_DUP_from_contigs.py:114-126 (asm10) and align_ins2ref.py:64-71 (map-*).  Here:

    recs = align_fastas(aligner, ref_fasta, [hp1_fasta, hp2_fasta], windows, out_sam, preset="asm5")
    pysam_like = AlignmentFile(out_sam)                   # .fetch(chrom) yields records with the fields the consumers read

Every contig is aligned against the reference window of the region it was assembled from (FocalSV assembles region by
region, so the window is known: `windows` maps a contig name to (chrom, start, end), or is parsed from the FASTA header
">name chrom:start-end"); minimap2 finds that window by seeding against the whole chromosome (row f2).  The records carry
what the consumers read: reference_name, pos / reference_start, reference_end, cigar / cigartuples (with clips),
qname / query_name, is_reverse, mapq / mapping_quality, seq / query_sequence, and the NM / AS tags (svim-asm reads NM,
software/svim-asm-1.0.2/src/svim_asm/SVIM_COLLECT.py:8-54).  The `--cs` tag is requested by the reference's command
line but never read and is not produced.  Output is plain SAM, coordinate-sorted (what
`samtools sort` would give); `samtools view -b` turns it into the BAM file the reference names.
"""
import numpy as np

from .hook import encode, realign_regions, realign_regions_chained

_OPS = "MIDNSHP=XB"


def read_fasta(path):
    """[(name, description, sequence)] of a FASTA file (multi-line records; name = header up to the first blank,
    utils.load_contigs' convention, focalsv/4_sv_calling/Dippav/utils.py:4-26)."""
    out, name, desc, parts = [], None, "", []
    with open(path) as fh:
        for line in fh:
            line = line.rstrip("\n").rstrip("\r")
            if line.startswith(">"):
                if name is not None:
                    out.append((name, desc, "".join(parts)))
                hdr = line[1:].split(None, 1)
                name, desc, parts = (hdr[0] if hdr else ""), (hdr[1] if len(hdr) > 1 else ""), []
            elif line and name is not None:
                parts.append(line)
    if name is not None:
        out.append((name, desc, "".join(parts)))
    return out


def reformat_contigs(records, hp):
    """reformat_fasta (DipPAV_variant_call.py:14-24): the n-th record of a haplotype FASTA is renamed contig_<hp>_<n>."""
    return [("contig_%s_%d" % (hp, k), desc, seq) for k, (_, desc, seq) in enumerate(records)]


def parse_window(desc):
    """'chr21:1000-21000' (anywhere in a FASTA description) -> ('chr21', 1000, 21000), 0-based half-open; None if absent."""
    for tok in desc.replace(",", " ").split():
        if ":" in tok and "-" in tok.split(":")[-1]:
            chrom, span = tok.rsplit(":", 1)
            a, b = span.split("-", 1)
            if a.isdigit() and b.isdigit():
                return chrom, int(a), int(b)
    return None


class SamRecord(object):
    """One alignment with pysam.AlignedSegment's names for the fields the reference reads."""
    __slots__ = ("qname", "flag", "reference_name", "pos", "mapq", "cigar", "seq", "tags", "reference_end")

    def __init__(self, qname, flag, reference_name, pos, mapq, cigar, seq, tags=None):
        self.qname, self.flag, self.reference_name, self.pos, self.mapq = qname, int(flag), reference_name, int(pos), int(mapq)
        self.cigar = [(int(op), int(n)) for op, n in cigar]
        self.seq = seq
        self.tags = dict(tags or {})
        self.reference_end = self.pos + sum(n for op, n in self.cigar if op in (0, 2, 3, 7, 8))

    # pysam's newer names
    query_name = property(lambda self: self.qname)
    reference_start = property(lambda self: self.pos)
    mapping_quality = property(lambda self: self.mapq)
    cigartuples = property(lambda self: self.cigar)
    query_sequence = property(lambda self: self.seq)
    is_reverse = property(lambda self: bool(self.flag & 16))
    is_supplementary = property(lambda self: bool(self.flag & 2048))
    is_secondary = property(lambda self: bool(self.flag & 256))
    is_unmapped = property(lambda self: bool(self.flag & 4))
    cigarstring = property(lambda self: "".join("%d%s" % (n, _OPS[op]) for op, n in self.cigar) or "*")
    query_length = property(lambda self: sum(n for op, n in self.cigar if op in (0, 1, 4, 7, 8)))

    def has_tag(self, t):
        return t in self.tags

    def get_tag(self, t):
        return self.tags[t]

    def to_sam(self):
        tags = "".join("\t%s:%s:%s" % (k, "i" if isinstance(v, (int, np.integer)) else "Z", v) for k, v in self.tags.items())
        return "%s\t%d\t%s\t%d\t%d\t%s\t*\t0\t0\t%s\t*%s" % (self.qname, self.flag, self.reference_name, self.pos + 1, self.mapq,
                                                           self.cigarstring, self.seq if self.seq else "*", tags)


def edit_distance_tag(cigar, q_codes, t_codes):
    """NM: mismatching M columns + inserted + deleted bases (SAM spec); q/t are code arrays of the aligned spans."""
    nm = qi = ti = 0
    for op, n in cigar:
        if op == 0:
            nm += int((q_codes[qi:qi + n] != t_codes[ti:ti + n]).sum()); qi += n; ti += n
        elif op == 1:
            nm += n; qi += n
        elif op == 2:
            nm += n; ti += n
        elif op == 4:
            qi += n
    return nm


def sam_records(aligned, windows, contigs):
    """hook.AlignedContig records -> SamRecord with sequence, NM and AS.  `windows`/`contigs` are what was aligned."""
    out = []
    for rec, (_, start, tseq), (_, qseq) in zip(aligned, windows, contigs):
        nm = edit_distance_tag(rec.cigar, encode(qseq), encode(tseq))
        seq = qseq if isinstance(qseq, str) else "".join("ACGTN"[int(c)] for c in qseq)
        out.append(SamRecord(rec.qname, 16 if rec.is_reverse else 0, rec.reference_name, rec.pos, rec.mapq, rec.cigar, seq,
                             {"NM": nm, "AS": int(rec.score)}))
    return out


def write_sam(path, records, ref_lengths, program="focalsv_b200", command=""):
    """Coordinate-sorted SAM (the order `samtools sort` gives: reference order of the header, then position; stable)."""
    order = {name: k for k, (name, _) in enumerate(ref_lengths)}
    recs = sorted(records, key=lambda r: (order[r.reference_name], r.pos))
    with open(path, "w") as fh:
        fh.write("@HD\tVN:1.6\tSO:coordinate\n")
        for name, ln in ref_lengths:
            fh.write("@SQ\tSN:%s\tLN:%d\n" % (name, ln))
        fh.write("@PG\tID:%s\tPN:%s\tCL:%s\n" % (program, program, command))
        for r in recs:
            fh.write(r.to_sam() + "\n")
    return recs


class AlignmentFile(object):
    """pysam.AlignmentFile for the two calls the reference makes on it: AlignmentFile(path) and .fetch(chrom) in coordinate
    order (extract_contig_signature_CCS.py:343-345).  Reads the SAM text write_sam produced (or takes records directly)."""

    def __init__(self, source, mode="r"):
        if isinstance(source, (list, tuple)):
            self._recs = list(source)
            self.references = tuple(dict.fromkeys(r.reference_name for r in self._recs))
            return
        self._recs, refs = [], []
        with open(source) as fh:
            for line in fh:
                if line.startswith("@"):
                    if line.startswith("@SQ"):
                        refs.append([f[3:] for f in line.rstrip("\n").split("\t") if f.startswith("SN:")][0])
                    continue
                f = line.rstrip("\n").split("\t")
                cigar, num = [], ""
                if f[5] != "*":
                    for ch in f[5]:
                        if ch.isdigit():
                            num += ch
                        else:
                            cigar.append((_OPS.index(ch), int(num))); num = ""
                tags = {}
                for t in f[11:]:
                    k, ty, v = t.split(":", 2)
                    tags[k] = int(v) if ty == "i" else v
                self._recs.append(SamRecord(f[0], int(f[1]), f[2], int(f[3]) - 1, int(f[4]), cigar, None if f[9] == "*" else f[9], tags))
        self.references = tuple(refs)

    def fetch(self, contig=None, start=None, stop=None):
        for r in sorted((r for r in self._recs if contig is None or r.reference_name == contig), key=lambda r: r.pos):
            if (start is None or r.reference_end > start) and (stop is None or r.pos < stop):
                yield r

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def align_fastas(aligner, ref_fasta, contig_fastas, windows=None, out_sam=None, preset="asm5", bw=2000, chained=False,
                 haplotypes=("hp1", "hp2")):
    """The minimap2 step of DipPAV_variant_call.py:83-112 on the GPU.

    ref_fasta: reference FASTA (one or more chromosomes).  contig_fastas: one FASTA per haplotype, renamed
    contig_hp1_<n> / contig_hp2_<n> as reformat_fasta does (pass haplotypes=None to keep the names).  windows: {original
    or renamed contig name: (chrom, start, end)}; a contig without an entry takes its window from its FASTA description
    ("chr21:1000-21000") and, failing that, is aligned against its whole chromosome when the reference has only one.
    chained=True decomposes every pair at minimizer-chain anchors first (hook.realign_regions_chained, row f2).
    Returns the SamRecord list in coordinate order; writes it to `out_sam` when given."""
    ref = read_fasta(ref_fasta)
    ref_by_name = {n: s for n, _, s in ref}
    pairs = []          # (qname, seq, chrom, start, end)
    for k, path in enumerate(contig_fastas):
        recs = read_fasta(path)
        named = reformat_contigs(recs, haplotypes[k]) if haplotypes else recs
        for (orig, desc, _), (name, _, seq) in zip(recs, named):
            win = (windows or {}).get(name) or (windows or {}).get(orig) or parse_window(desc)
            if win is None:
                if len(ref) != 1:
                    raise ValueError("contig %s: no window given and the reference has %d sequences" % (name, len(ref)))
                win = (ref[0][0], 0, len(ref[0][2]))
            chrom, s, e = win
            if chrom not in ref_by_name or not (0 <= s < e <= len(ref_by_name[chrom])):
                raise ValueError("contig %s: window %s:%d-%d is outside the reference" % (name, chrom, s, e))
            pairs.append((name, seq, chrom, int(s), int(e)))
    wins = [(chrom, s, ref_by_name[chrom][s:e]) for _, _, chrom, s, e in pairs]
    tigs = [(name, seq) for name, seq, _, _, _ in pairs]
    fn = realign_regions_chained if chained else realign_regions
    aligned = fn(aligner, wins, tigs, preset=preset, bw=bw)
    recs = sam_records(aligned, wins, tigs)
    ref_lengths = [(n, len(s)) for n, _, s in ref]
    if out_sam:
        return write_sam(out_sam, recs, ref_lengths, command="align_fastas -x %s -r%d %s %s" % (preset, bw, ref_fasta, " ".join(contig_fastas)))
    order = {name: k for k, (name, _) in enumerate(ref_lengths)}
    return sorted(recs, key=lambda r: (order[r.reference_name], r.pos))
