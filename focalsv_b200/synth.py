"""Synthetic workloads shaped like BASELINE.json's five configs.

No network, no real genomes: references are i.i.d. ACGT (GC 0.41), region
lengths are resampled from the quantiles of the BED files the reference ships
(test/SV_Regions_HG002_HIFI_L1_FocalSV-auto.bed, target_sv/HCC1395_SV_rich_regions_DUP.bed),
planted SV sizes from the GIAB chr21 truth set shipped in
focalsv/backup/test/test_targetSV_chr21.vcf.  Everything is
seeded (1001..1004) and deterministic.

A workload is a list of `Group`s; one group = one scoring = one fsv_align_batch call.
"""
from collections import namedtuple

import numpy as np

from . import _abi
from .presets import PRESETS, ksw_band, scoring_for

Group = namedtuple("Group", "name preset scoring qarena tarena tasks region_of")

# quantile tables (probability, value); interpolated in log space
REGION_LEN_Q = ((0.0, 14000), (0.5, 14233), (0.9, 49342), (0.99, 120675), (1.0, 1146440))
DUP_LEN_Q = ((0.0, 100052), (0.5, 110707), (0.9, 323372), (0.99, 4211011), (1.0, 4824525))
SV_LEN_Q = ((0.0, 50), (0.5, 168), (0.9, 1311), (1.0, 12599))


def sample_quantiles(rng, table, n, cap=None):
    p = np.array([t[0] for t in table])
    v = np.log(np.array([float(t[1]) for t in table]))
    x = np.exp(np.interp(rng.random(n), p, v))
    if cap is not None:
        x = np.minimum(x, cap)
    return np.maximum(x.astype(np.int64), 1)


def random_seq(rng, n, gc=0.41):
    r = rng.random(n)
    at, g = (1.0 - gc) / 2.0, gc / 2.0
    out = np.zeros(n, dtype=np.uint8)          # A
    out[r >= at] = 1                           # C
    out[r >= at + g] = 2                       # G
    out[r >= at + 2 * g] = 3                   # T
    return out


def mutate(rng, seq, sub=0.0, ins=0.0, dele=0.0):
    """Per-base substitution / 1-bp insertion (before the base) / deletion."""
    n = len(seq)
    if n == 0 or (sub + ins + dele) <= 0:
        return seq.copy()
    r = rng.random(n)
    is_sub = r < sub
    is_ins = (r >= sub) & (r < sub + ins)
    is_del = (r >= sub + ins) & (r < sub + ins + dele)
    base = seq.copy()
    k = int(is_sub.sum())
    if k:
        base[is_sub] = (base[is_sub] + 1 + rng.integers(0, 3, k).astype(np.uint8)) & 3
    counts = np.ones(n, dtype=np.int64)
    counts[is_del] = 0
    counts[is_ins] = 2
    out = np.repeat(base, counts)
    k = int(is_ins.sum())
    if k:
        starts = (np.cumsum(counts) - counts)[is_ins]
        out[starts] = rng.integers(0, 4, k).astype(np.uint8)
    return out


def plant_svs(rng, ref, n_sv, max_net=1300, max_len=None, sizes=None):
    """Haplotype contig = reference window with n_sv planted DEL/INS.

    The running length difference is steered back towards zero so that a global
    alignment stays inside a band of half-width `max_net`.  Returns (contig, svs)
    with svs = [(ref_pos, 'DEL'|'INS', length)]."""
    n = len(ref)
    if n_sv <= 0 or n < 2000:
        return ref.copy(), []
    pos = np.sort(rng.integers(500, n - 500, n_sv))
    if sizes is None:
        sizes = sample_quantiles(rng, SV_LEN_Q, n_sv, cap=max_len)
    parts, svs, last, net = [], [], 0, 0
    for p, L in zip(pos, sizes):
        p, L = int(p), int(L)
        if p < last:
            continue
        is_del = rng.random() < 0.5
        if net - L < -max_net:
            is_del = False
        if net + L > max_net:
            is_del = True
        if is_del and net - L < -max_net:
            continue
        parts.append(ref[last:p])
        if is_del:
            if p + L >= n - 100:
                last = p
                continue
            last = p + L
            net -= L
            svs.append((p, "DEL", L))
        else:
            parts.append(random_seq(rng, L))
            last = p
            net += L
            svs.append((p, "INS", L))
    parts.append(ref[last:])
    return np.concatenate(parts), svs


def _pack(name, preset, pairs, w, zdrop, flag=0, end_bonus=0, regions=None, flags=None):
    """pairs: list of (query, target) uint8 arrays -> Group."""
    qlens = [len(q) for q, _ in pairs]
    tlens = [len(t) for _, t in pairs]
    qarena = np.concatenate([q for q, _ in pairs]) if pairs else np.zeros(0, np.uint8)
    tarena = np.concatenate([t for _, t in pairs]) if pairs else np.zeros(0, np.uint8)
    tasks = np.zeros(len(pairs), dtype=_abi.TASK_DTYPE)
    tasks["qlen"], tasks["tlen"] = qlens, tlens
    tasks["q_off"] = np.concatenate([[0], np.cumsum(qlens)[:-1]]) if pairs else []
    tasks["t_off"] = np.concatenate([[0], np.cumsum(tlens)[:-1]]) if pairs else []
    tasks["w"], tasks["zdrop"], tasks["end_bonus"] = w, zdrop, end_bonus
    tasks["flag"] = flag if flags is None else flags
    region_of = np.arange(len(pairs), dtype=np.int64) if regions is None else np.asarray(regions, dtype=np.int64)
    return Group(name, preset, scoring_for(preset), qarena, tarena, tasks, region_of)


def contig_pairs(rng, region_lens, sv_per_bp, err, max_net, max_sv=None, n_hap=2):
    pairs, regions = [], []
    for ri, L in enumerate(region_lens):
        L = int(L)
        ref = random_seq(rng, L)
        shared_seed = int(rng.integers(1 << 30))
        for h in range(n_hap):
            # hp1 and hp2 share about half of their SVs: same sub-stream for the shared half
            n_sv = max(int(rng.poisson(L * sv_per_bp)), 0)
            hrng = np.random.default_rng(shared_seed + (0 if rng.random() < 0.5 else h + 1))
            contig, _ = plant_svs(hrng, ref, n_sv, max_net=max_net, max_len=max_sv)
            contig = mutate(rng, contig, sub=err * 0.6, ins=err * 0.2, dele=err * 0.2) if err > 0 else contig
            pairs.append((contig, ref))
            regions.append(ri)
    return pairs, regions


def config2(n_regions=5000, seed=1002, max_region=None, lens=None):
    """cfg2: auto mode, ~5k SV-rich regions, HiFi contigs vs hg38-shaped reference, asm5, -r2k.
    `lens` overrides the sampled region lengths (used by the multi-GPU sharding in bench.py)."""
    rng = np.random.default_rng(seed)
    if lens is None:
        lens = sample_quantiles(rng, REGION_LEN_Q, n_regions, cap=max_region)
    p = PRESETS["asm5"]
    w = ksw_band(p.bw)
    pairs, regions = contig_pairs(rng, lens, 1.0 / 15000.0, 0.001, max_net=w // 2 - 200, max_sv=w // 2 - 300)
    return [_pack("cfg2.contigs.asm5", "asm5", pairs, w, p.zdrop, flag=0, regions=regions)]


def config1(seed=1001, window=200000, n_reads=400):
    """cfg1: target mode, one 200 kb chr21 window: 2 contigs (asm5) + 400 HiFi reads (map-hifi)."""
    rng = np.random.default_rng(seed)
    ref = random_seq(rng, window)
    pa = PRESETS["asm5"]
    w = ksw_band(pa.bw)
    pairs = []
    for h in range(2):
        sizes = sample_quantiles(rng, SV_LEN_Q, 10, cap=w // 2 - 300)
        contig, _ = plant_svs(rng, ref, 10, max_net=w // 2 - 200, sizes=sizes)
        pairs.append((mutate(rng, contig, sub=0.0006, ins=0.0002, dele=0.0002), ref))
    g1 = _pack("cfg1.contigs.asm5", "asm5", pairs, w, pa.zdrop, regions=[0, 0])
    ph = PRESETS["map-hifi"]
    wr = ksw_band(500)
    rp = []
    for _ in range(n_reads):
        L = int(rng.integers(15000, 20001))
        s = int(rng.integers(0, window - L))
        rp.append((mutate(rng, ref[s:s + L], sub=0.003, ins=0.001, dele=0.001), ref[s:s + L]))
    g2 = _pack("cfg1.reads.map-hifi", "map-hifi", rp, wr, ph.zdrop, regions=[0] * n_reads)
    return [g1, g2]


def config3(n_regions=2000, seed=1003, max_region=None):
    """cfg3: ONT reads/contigs, wide band w=500 with z-drop; single-affine (in-tree constants) and map-ont."""
    rng = np.random.default_rng(seed)
    lens = sample_quantiles(rng, REGION_LEN_Q, n_regions, cap=max_region)
    groups = []
    # contigs with 1% error vs window, single-affine a=2,b=4,q=4,e=2, w=500, z=400 (Correct.h:1194-1199)
    pairs, regions = contig_pairs(rng, lens, 1.0 / 15000.0, 0.01, max_net=200, max_sv=None)
    flags = np.where(np.arange(len(pairs)) % 2 == 0, 0, _abi.EZ_EXTZ_ONLY).astype(np.int32)
    groups.append(_pack("cfg3.contigs.extz2", "hifiasm", pairs, 500, 400, regions=regions, flags=flags))
    # ONT reads 10-50 kb, 4% sub / 3% ins / 3% del, map-ont dual affine
    rp, rr = [], []
    for ri, L in enumerate(lens):
        L = int(L)
        ref = random_seq(rng, L)
        rl = int(min(L, rng.integers(10000, 50001)))
        s = int(rng.integers(0, L - rl + 1))
        rp.append((mutate(rng, ref[s:s + rl], sub=0.04, ins=0.03, dele=0.03), ref[s:s + rl]))
        rr.append(ri)
    flags = np.where(np.arange(len(rp)) % 2 == 0, 0, _abi.EZ_EXTZ_ONLY).astype(np.int32)
    groups.append(_pack("cfg3.reads.map-ont", "map-ont", rp, 500, 400, regions=rr, flags=flags))
    return groups


def config4(seed=1004, n_dup=213, n_pair=270, max_region=None):
    """cfg4: HCC1395-shaped TRA/INV/DUP realignment: CLR contig ends (asm10, extension-only) and
    INS alleles vs +-(len+2 kb) windows (map-pb)."""
    rng = np.random.default_rng(seed)
    pa = PRESETS["asm10"]
    w = ksw_band(pa.bw)
    lens = sample_quantiles(rng, DUP_LEN_Q, n_dup, cap=max_region)
    pairs, regions = [], []
    for ri, L in enumerate(lens):
        L = int(L)
        ref = random_seq(rng, L)
        # tandem duplication junction: contig follows the window, then jumps back
        cut = int(rng.integers(L // 2, L - 1000))
        dup = int(min(rng.integers(1000, 50000), cut - 500))
        contig = np.concatenate([ref[:cut], ref[cut - dup:]])
        contig = mutate(rng, contig, sub=0.006, ins=0.002, dele=0.002)
        pairs.append((contig, ref)); regions.append(ri)                       # prefix extension
        pairs.append((contig[::-1].copy(), ref[::-1].copy())); regions.append(ri)  # suffix extension (reversed)
    for k in range(n_pair):
        a, b = random_seq(rng, 100000), random_seq(rng, 100000)
        cut = int(rng.integers(20000, 80000))
        contig = mutate(rng, np.concatenate([a[:cut], b[cut:]]), sub=0.006, ins=0.002, dele=0.002)
        pairs.append((contig, a)); regions.append(n_dup + k)
        pairs.append((contig[::-1].copy(), b[::-1].copy())); regions.append(n_dup + k)
    g1 = _pack("cfg4.contig-ends.asm10", "asm10", pairs, w, pa.zdrop, flag=_abi.EZ_EXTZ_ONLY, regions=regions)
    pp = PRESETS["map-pb"]
    ip, ir = [], []
    sizes = sample_quantiles(rng, SV_LEN_Q, n_dup)
    for ri, L in enumerate(sizes):
        L = int(L)
        win = random_seq(rng, 2 * L + 4000)
        allele = mutate(rng, win[L + 2000 - L:L + 2000], sub=0.006, ins=0.002, dele=0.002)
        ip.append((allele, win)); ir.append(ri)
    g2 = _pack("cfg4.ins-alleles.map-pb", "map-pb", ip, w, pp.zdrop, flag=_abi.EZ_EXTZ_ONLY, regions=ir)
    return [g1, g2]


def config5(**kw):
    """cfg5 = the cfg2 batch, sharded across 1/2/4/8 GPUs."""
    return config2(**kw)


CONFIGS = {"cfg1": config1, "cfg2": config2, "cfg3": config3, "cfg4": config4, "cfg5": config5}


def small_mixed(seed=7, n=24, max_len=1200):
    """Small parity workload: mixed lengths / bands / flags, one group per gap model."""
    rng = np.random.default_rng(seed)
    groups = []
    for preset, w in (("hifiasm", 100), ("asm5", 151), ("map-ont", 75)):
        pairs, flags = [], []
        for i in range(n):
            L = int(rng.integers(30, max_len))
            ref = random_seq(rng, L)
            q, _ = plant_svs(rng, ref, int(rng.integers(0, 3)), max_net=w // 2 - 10, max_len=w // 2 - 10) if L > 2000 else (ref.copy(), [])
            q = mutate(rng, q, sub=0.02, ins=0.01, dele=0.01)
            if len(q) == 0:
                q = ref[:1].copy()
            pairs.append((q, ref))
            flags.append([0, _abi.EZ_EXTZ_ONLY, _abi.EZ_RIGHT, _abi.EZ_EXTZ_ONLY | _abi.EZ_REV_CIGAR | _abi.EZ_RIGHT][i % 4])
        p = PRESETS[preset]
        groups.append(_pack("small." + preset, preset, pairs, w, p.zdrop, flags=np.array(flags, dtype=np.int32)))
    return groups
