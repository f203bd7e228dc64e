"""Seeded inputs of the signature -> VCF tests (shared by tests/golden/make_vcf_golden.py, which feeds them to the
REFERENCE'S OWN functions, and tests/test_sv_call.py, which feeds them to focalsv_b200.sv_call)."""
import numpy as np

from focalsv_b200 import synth
from focalsv_b200.dropin import SamRecord

PLATFORM_ERR = {"CCS": 0.001, "CLR": 0.003, "ONT": 0.003}
CHROM = "chr21"


def ascii_of(codes):
    return "".join("ACGTN"[int(c)] for c in codes)


def region_case(seed, platform, n_regions=8):
    """A scaled config: one chromosome, n_regions target regions, an hp1 and an hp2 contig per region (about half of
    the planted SVs shared between the haplotypes, as synth.contig_pairs plants them).
    Returns (chrom_codes, windows[(chrom, start, codes)], contigs[(qname, codes)])."""
    rng = np.random.default_rng(seed)
    lens = [int(x) for x in rng.integers(9000, 22000, n_regions)]
    gap = 3000
    chrom = synth.random_seq(rng, sum(lens) + gap * (n_regions + 1))
    windows, contigs = [], []
    pos = gap
    err = PLATFORM_ERR[platform]
    for i, L in enumerate(lens):
        ref = chrom[pos:pos + L]
        shared = int(rng.integers(1 << 30))
        n_shared, n_private = int(rng.integers(1, 4)), [int(rng.integers(0, 3)), int(rng.integers(0, 3))]
        for h in (1, 2):
            # homozygous SVs: the same stream plants them into both haplotypes; then each haplotype gets its own
            q, _ = synth.plant_svs(np.random.default_rng(shared), ref, n_shared, max_net=700, max_len=650)
            q, _ = synth.plant_svs(np.random.default_rng(shared + h), q, n_private[h - 1], max_net=500, max_len=450)
            q = synth.mutate(rng, q, err * 0.6, err * 0.2, err * 0.2)
            windows.append((CHROM, pos, ref)); contigs.append(("contig_hp%d_%d" % (h, i), q))
        pos += L + gap
    return chrom, windows, contigs


def _random_cigar(rng, dense):
    cig = []
    n = int(rng.integers(2, 30))
    for i in range(n):
        cig.append((0, int(rng.integers(1, 300 if dense or rng.random() < 0.7 else 4000))))
        if i + 1 < n:
            op = 1 if rng.random() < 0.5 else 2
            cig.append((op, int(rng.integers(25, 1200)) if rng.random() < 0.6 else int(rng.integers(1, 40))))
            if rng.random() < 0.15:
                cig.append((3 - op, int(rng.integers(20, 600))))
    return cig


def record_case(seed, n_contigs=40, chrom_len=300000):
    """Random alignment records on one chromosome: single records with dense DEL/INS (clusters, pairs, the CLR record
    filter) and contigs split into two or three clipped records (every branch of extract_sig_from_split: deletions with and
    without reference overlap, insertions, strand / mapq filters).  Returns (chrom_codes, {qname: contig_ascii}, [SamRecord])."""
    rng = np.random.default_rng(seed)
    chrom = synth.random_seq(rng, chrom_len)
    recs, tigs = [], {}
    for k in range(n_contigs):
        hp = 1 + k % 2
        name = "contig_hp%d_%d" % (hp, k // 2)
        mapq = int(rng.choice([60, 60, 60, 50, 20, 0]))
        rev = bool(rng.random() < 0.25)
        # hp2 contigs often sit near their hp1 twin so that pair_sig has something to pair
        if hp == 2 and rng.random() < 0.7 and recs:
            base = max(0, recs[-1].pos + int(rng.integers(-150, 150)))
        else:
            base = int(rng.integers(0, chrom_len - 60000))
        if rng.random() < 0.45:
            # a contig in 2-3 pieces: [M.. S] [S M.. S] [S M..]
            n_piece = 2 if rng.random() < 0.7 else 3
            body = [_random_cigar(rng, dense=False) if rng.random() < 0.5 else [(0, int(rng.integers(300, 3000)))] for _ in range(n_piece)]
            qlens = [sum(n for op, n in c if op in (0, 1)) for c in body]
            extra = [int(rng.choice([0, 0, 40, 300, 1500])) for _ in range(n_piece - 1)]         # contig bases between the pieces (an insertion)
            total = sum(qlens) + sum(extra)
            tigs[name] = ascii_of(synth.random_seq(rng, total))
            qpos, rpos = 0, base
            for pi in range(n_piece):
                lead, trail = qpos, total - qpos - qlens[pi]
                clip = 5 if rng.random() < 0.3 else 4
                cig = ([(clip, lead)] if lead else []) + body[pi] + ([(clip, trail)] if trail else [])
                r = SamRecord(name, (16 if rev else 0) | (2048 if pi else 0), CHROM, min(rpos, chrom_len - 40000), mapq if rng.random() < 0.9 else 10, cig, None)
                recs.append(r)
                qpos += qlens[pi] + (extra[pi] if pi < n_piece - 1 else 0)
                # next piece: a reference gap (deletion), an overlap, or flush
                jump = int(rng.choice([0, 0, 35, 200, 2500, -20, -500, -2500, 60000]))
                rpos = max(0, r.reference_end + jump)
        else:
            cig = _random_cigar(rng, dense=bool(rng.random() < 0.3))
            if rng.random() < 0.2:
                cig = [(5 if rng.random() < 0.5 else 4, int(rng.integers(1, 2000)))] + cig
            tigs[name] = ascii_of(synth.random_seq(rng, sum(n for op, n in cig if op in (0, 1, 4, 5))))
            recs.append(SamRecord(name, 16 if rev else 0, CHROM, base, mapq, cig, None))
    recs = [r for r in recs if r.reference_end < chrom_len]
    return chrom, tigs, recs


CASES = [("regions", "CCS", 11), ("regions", "CLR", 12), ("regions", "ONT", 13),
         ("records", "CCS", 21), ("records", "CLR", 22), ("records", "ONT", 23), ("records", "CCS", 24), ("records", "ONT", 25)]
