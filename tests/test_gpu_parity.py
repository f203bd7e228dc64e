"""GPU parity proper: libfocalsv_cuda (through the C ABI) vs the CPU oracle, bit-exact.

Integer/byte/index work: every ksw_extz_t field, every CIGAR word and the in-band
cell count must be identical (no tolerance)."""
import numpy as np
import pytest

from focalsv_b200 import _abi, synth
from focalsv_b200.api import FsvError, make_tasks, task_cigar
from util import compare_group, describe, random_case, same_result

pytestmark = pytest.mark.gpu


def _fuzz(O, al, seed, n, dual, max_len=400):
    rng = np.random.default_rng(seed)
    bad = []
    for it in range(n):
        c = random_case(rng, max_len=max_len, dual=dual)
        fo = O.extd2 if dual else O.extz2
        fg = al.extd2 if dual else al.extz2
        r1, c1 = fo(c["q"], c["t"], c["sc"], w=c["w"], zdrop=c["zdrop"], end_bonus=c["end_bonus"], flag=c["flag"])
        r2, c2 = fg(c["q"], c["t"], c["sc"], w=c["w"], zdrop=c["zdrop"], end_bonus=c["end_bonus"], flag=c["flag"])
        if not same_result(r1, c1, r2, c2) or int(r1["cells"]) != int(r2["cells"]):
            bad.append((it, len(c["q"]), len(c["t"]), c["w"], c["zdrop"], hex(c["flag"]), describe(r1, c1), describe(r2, c2)))
    return bad


def test_fuzz_single_affine_exact_kernel(oracle, aligner):
    aligner.set_option("force_exact", 1)
    try:
        bad = _fuzz(oracle, aligner, 11, 250, dual=False)
    finally:
        aligner.set_option("force_exact", 0)
    assert not bad, bad[:3]


def test_fuzz_dual_affine_exact_kernel(oracle, aligner):
    aligner.set_option("force_exact", 1)
    try:
        bad = _fuzz(oracle, aligner, 12, 250, dual=True)
    finally:
        aligner.set_option("force_exact", 0)
    assert not bad, bad[:3]


def test_fuzz_single_affine_default_routing(oracle, aligner):
    bad = _fuzz(oracle, aligner, 13, 250, dual=False)
    assert not bad, bad[:3]


def test_fuzz_dual_affine_default_routing(oracle, aligner):
    bad = _fuzz(oracle, aligner, 14, 250, dual=True)
    assert not bad, bad[:3]


@pytest.mark.parametrize("force_exact", [0, 1])
def test_small_mixed_batches(oracle, aligner, force_exact):
    aligner.set_option("force_exact", force_exact)
    try:
        for g in synth.small_mixed(seed=7, n=24):
            bad, ores, gres = compare_group(oracle, aligner, g)
            assert not bad, (g.name, bad[:5])
    finally:
        aligner.set_option("force_exact", 0)


def test_edge_cases(oracle, aligner):
    sc = _abi.make_scoring(2, 4, 4, 2, sc_ambi=0)
    sd = _abi.make_scoring(1, 19, 39, 3, 81, 1)
    one = np.array([2], dtype=np.uint8)
    seq = np.random.default_rng(3).integers(0, 4, 77).astype(np.uint8)
    for q, t in ((one, one), (one, seq), (seq, one), (seq, seq), (seq[:16], seq[:17]), (seq[:33], seq[:15])):
        for w in (-1, 0, 1, 8):
            for flag in (0, _abi.EZ_EXTZ_ONLY, _abi.EZ_SCORE_ONLY, _abi.EZ_RIGHT | _abi.EZ_REV_CIGAR):
                r1, c1 = oracle.extz2(q, t, sc, w=w, zdrop=50, flag=flag)
                r2, c2 = aligner.extz2(q, t, sc, w=w, zdrop=50, flag=flag)
                assert same_result(r1, c1, r2, c2), (len(q), len(t), w, flag, describe(r1, c1), describe(r2, c2))
                r1, c1 = oracle.extd2(q, t, sd, w=w, zdrop=50, flag=flag)
                r2, c2 = aligner.extd2(q, t, sd, w=w, zdrop=50, flag=flag)
                assert same_result(r1, c1, r2, c2), (len(q), len(t), w, flag, describe(r1, c1), describe(r2, c2))


def test_empty_and_reset_tasks(oracle, aligner):
    """ksw2's silent returns (ksw2_extz2_sse.c:57,82) come back as successful tasks with reset results."""
    sc = _abi.make_scoring(2, 4, 4, 2, sc_ambi=0)
    seq = np.arange(40, dtype=np.uint8) & 3
    tasks = make_tasks([0, 40, 40], [40, 0, 40], 50, 100)
    tasks["q_off"] = 0
    tasks["t_off"] = 0
    res, cig = aligner.align_batch(sc, seq, seq, tasks)
    for i in (0, 1):
        assert int(res[i]["score"]) == _abi.NEG_INF and int(res[i]["max"]) == 0 and int(res[i]["n_cigar"]) == 0
        assert int(res[i]["max_t"]) == -1 and int(res[i]["mqe_t"]) == -1
    assert _abi.cigar_str(task_cigar(res[2], cig)) == "40M"
    # mismatch score that int8 lanes cannot represent: -min_sc > 2(q+e)  (:82)
    bad_sc = _abi.make_scoring(1, 60, 4, 2, sc_ambi=0)
    res, cig = aligner.align_batch(bad_sc, seq, seq, tasks[2:])
    r1, _ = oracle.extz2(seq, seq, bad_sc, w=50, zdrop=100)
    assert int(res[0]["score"]) == _abi.NEG_INF == int(r1["score"]) and int(res[0]["n_cigar"]) == 0
    res, cig = aligner.align_batch(sc, seq, seq, tasks[:0])
    assert len(res) == 0 and len(cig) == 0


def test_cigar_arena_overflow_reports_required_size(aligner):
    sc = _abi.make_scoring(2, 4, 4, 2, sc_ambi=0)
    rng = np.random.default_rng(5)
    t = rng.integers(0, 4, 300).astype(np.uint8)
    q = synth.mutate(rng, t, 0.05, 0.03, 0.03)
    tasks = make_tasks([len(q)], [len(t)], -1, -1)
    with pytest.raises(FsvError) as ei:
        aligner.align_batch(sc, q, t, tasks, cigar_cap=2)
    assert ei.value.code == _abi.ERR_CIGAR_CAP


def test_bad_offsets_rejected(aligner):
    sc = _abi.make_scoring(2, 4, 4, 2, sc_ambi=0)
    seq = np.zeros(10, dtype=np.uint8)
    tasks = make_tasks([20], [10], -1, -1)
    with pytest.raises(FsvError) as ei:
        aligner.align_batch(sc, seq, seq, tasks)
    assert ei.value.code == _abi.ERR_INVALID


def test_staged_batch_equals_one_call(oracle, aligner):
    g = synth.small_mixed(seed=9, n=12)[1]
    res1, cig1 = aligner.align_batch(g.scoring, g.qarena, g.tarena, g.tasks)
    b = aligner.batch(g.scoring, g.qarena, g.tarena, g.tasks)
    b.run(); b.run()          # re-running a resident batch is idempotent
    res2, cig2 = b.fetch()
    b.close()
    for f in _abi.EZ_FIELDS:
        assert np.array_equal(res1[f], res2[f])
    for i in range(len(res1)):
        assert np.array_equal(task_cigar(res1[i], cig1), task_cigar(res2[i], cig2))


def test_medium_tasks_band_3001(oracle, aligner):
    """cfg2-shaped but short enough for the oracle: asm5, w=3001, 6-12 kb contigs with planted SVs."""
    rng = np.random.default_rng(21)
    pairs, regions = synth.contig_pairs(rng, [6000, 9000, 12000], 1.0 / 4000.0, 0.002, max_net=1300, max_sv=1200)
    g = synth._pack("medium.asm5", "asm5", pairs, 3001, 200, regions=regions)
    before = aligner.stats()["exact_path_tasks"]
    bad, ores, gres = compare_group(oracle, aligner, g)
    assert not bad, bad
    assert aligner.stats()["exact_path_tasks"] == before      # all of them ran on the DPX kernel


def _golden(name):
    import os
    from test_oracle_golden import load
    return load(name)


@pytest.mark.parametrize("fname,dual", [("kat_extz2.npz", False), ("kat_extd2.npz", True)])
def test_golden_vectors_on_gpu(aligner, fname, dual):
    """tests/golden: outputs of the reference's own compiled ksw_extz2_sse / frozen dual-affine vectors."""
    for c in _golden(fname):
        f = aligner.extd2 if dual else aligner.extz2
        r, cig = f(c["q"], c["t"], c["sc"], w=c["w"], zdrop=c["zdrop"], end_bonus=c["end_bonus"], flag=c["flag"])
        assert same_result(c["res"], c["cigar"], r, cig), (c["name"], describe(c["res"], c["cigar"]), describe(r, cig))


@pytest.mark.parametrize("preset,w,lens", [("asm5", 3001, (7000, 9000)), ("asm10", 3900, (8500, 9500)),
                                            ("map-hifi", 751, (3000, 5000)), ("hifiasm", 500, (2500, 4000)),
                                            ("map-ont", 2300, (5000, 6000))])
def test_wide_band_classes(oracle, aligner, preset, w, lens):
    """One task per DPX warps-per-task class (2, 4, 6, 8 warps) plus flags, against the oracle."""
    rng = np.random.default_rng(31 + w)
    pairs, flags = [], []
    for i, L in enumerate(lens):
        ref = synth.random_seq(rng, L)
        q, _ = synth.plant_svs(rng, ref, 3, max_net=min(w // 2 - 50, 1000), max_len=min(w // 2 - 60, 800))
        pairs.append((synth.mutate(rng, q, 0.004, 0.002, 0.002), ref))
        flags.append([0, _abi.EZ_EXTZ_ONLY][i % 2])
    from focalsv_b200.presets import PRESETS
    g = synth._pack("wide." + preset, preset, pairs, w, PRESETS[preset].zdrop, flags=np.array(flags, dtype=np.int32))
    before = aligner.stats()["exact_path_tasks"]
    bad, ores, gres = compare_group(oracle, aligner, g)
    assert not bad, (preset, bad)
    assert aligner.stats()["exact_path_tasks"] == before


def test_full_size_properties_cfg2_slice(oracle, aligner):
    """BASELINE-size tasks (too big for the oracle's full check inside the test budget): size-independent
    properties.  CIGAR consumes both sequences and re-scores to ez.score; score-only == with-CIGAR; re-running is
    idempotent; cells equal the band definition when the task is not dropped."""
    g = synth.config2(n_regions=12, seed=5, max_region=120000)[0]
    res, cig = aligner.align_batch(g.scoring, g.qarena, g.tarena, g.tasks)
    res2, cig2 = aligner.align_batch(g.scoring, g.qarena, g.tarena, g.tasks)
    so = g.tasks.copy()
    so["flag"] |= _abi.EZ_SCORE_ONLY
    res3, _ = aligner.align_batch(g.scoring, g.qarena, g.tarena, so)
    for f in _abi.EZ_FIELDS:
        assert np.array_equal(res[f], res2[f])
        if f != "n_cigar":
            assert np.array_equal(res[f], res3[f]), f
    from focalsv_b200 import api
    for i in range(len(res)):          # the arena order depends on completion order; per-task CIGARs do not
        assert np.array_equal(task_cigar(res[i], cig), task_cigar(res2[i], cig2))
    n_ok = 0
    for i, t in enumerate(g.tasks):
        q = g.qarena[t["q_off"]:t["q_off"] + t["qlen"]]
        tt = g.tarena[t["t_off"]:t["t_off"] + t["tlen"]]
        if int(res[i]["zdropped"]):
            continue
        n_ok += 1
        c = task_cigar(res[i], cig)
        s, qu, tu = oracle.score_cigar(q, tt, g.scoring, c)
        assert (qu, tu) == (int(t["qlen"]), int(t["tlen"]))
        assert s == int(res[i]["score"]), (i, s, int(res[i]["score"]))
        assert int(res[i]["cells"]) == api.task_cells(int(t["qlen"]), int(t["tlen"]), int(t["w"]))
    assert n_ok >= len(g.tasks) // 2


def test_small_pages_and_tight_pool(oracle):
    """Traceback paging: 64 KB pages (every task spans many pages) and a pool that only fits a few tasks at a
    time (CTAs must wait for pages and take small tasks meanwhile) give the same bits."""
    from focalsv_b200 import api
    al = api.Aligner(0)
    try:
        al.set_option("traceback_page_bytes", 65536)
        al.set_option("traceback_budget_bytes", 120 * 65536)
        rng = np.random.default_rng(123)
        pairs = []
        for L in list(rng.integers(200, 3000, 60)) + [6000, 7000]:
            ref = synth.random_seq(rng, int(L))
            pairs.append((synth.mutate(rng, ref, 0.01, 0.005, 0.005), ref))
        from focalsv_b200.presets import PRESETS
        # lazy_min_pages: 16 = default (the longest tasks take their pages as they advance, the others up front),
        # 3 = nearly every task grows lazily (banker's check on every grant), 0 = every task up front
        for lazy in (16, 3, 0):
            al.set_option("lazy_min_pages", lazy)
            al.set_option("pool_stall_ms", 5000)
            for preset, w in (("asm5", 301), ("hifiasm", 100)):
                g = synth._pack("paged." + preset, preset, pairs, w, PRESETS[preset].zdrop)
                bad, ores, gres = compare_group(oracle, al, g)
                assert not bad, (lazy, preset, bad[:5])
        # a task whose traceback cannot fit the pool at all is refused, not hung
        ref = synth.random_seq(rng, 40000)
        g = synth._pack("toolarge", "asm5", [(ref.copy(), ref)], 3001, 200)
        with pytest.raises(FsvError) as ei:
            al.align_batch(g.scoring, g.qarena, g.tarena, g.tasks)
        assert ei.value.code == _abi.ERR_NOMEM
    finally:
        al.close()


@pytest.mark.parametrize("preset,w", [("asm5", 3001), ("hifiasm", 500), ("map-ont", 40)])
def test_wildcard_bases_on_the_dpx_kernel(oracle, aligner, preset, w):
    """N bases (code 4, scored sc_N, ksw2_extz2_sse.c:68,130-134) in the query, the target or both: isolated,
    in runs that cross 16-lane vectors and warps, and at the ends; the tasks stay on the DPX kernel."""
    rng = np.random.default_rng(77 + w)
    pairs = []
    for i, L in enumerate([700, 1500, 2600, 4100, 5200, 6400]):
        ref = synth.random_seq(rng, L)
        q = synth.mutate(rng, ref, 0.01, 0.004, 0.004).copy()
        ref = ref.copy()
        if i % 3 != 1:
            q[rng.integers(0, len(q), len(q) // 30)] = 4
            k = int(rng.integers(0, len(q) - 200)); q[k:k + int(rng.integers(1, 150))] = 4
        if i % 3 != 0:
            ref[rng.integers(0, len(ref), len(ref) // 40)] = 4
            k = int(rng.integers(0, len(ref) - 700)); ref[k:k + int(rng.integers(1, 600))] = 4
        if i == 4:
            q[:3] = 4; q[-2:] = 4; ref[0] = 4; ref[-1] = 4
        pairs.append((q, ref))
    flags = np.array([0, _abi.EZ_EXTZ_ONLY, 0, _abi.EZ_SCORE_ONLY, _abi.EZ_REV_CIGAR, 0], dtype=np.int32)
    from focalsv_b200.presets import PRESETS
    g = synth._pack("wild." + preset, preset, pairs, w, PRESETS[preset].zdrop, flags=flags)
    before = aligner.stats()["exact_path_tasks"]
    bad, ores, gres = compare_group(oracle, aligner, g)
    assert not bad, bad
    assert aligner.stats()["exact_path_tasks"] == before


@pytest.mark.parametrize("preset,w", [("asm5", 3001), ("hifiasm", 500), ("map-hifi", 751)])
def test_right_aligned_gaps_on_the_dpx_kernel(oracle, aligner, preset, w):
    """KSW_EZ_RIGHT (ties go to the later state, continuation bits on >= 0; ksw2_extz2_sse.c:197-222) as minimap2's
    left extension uses it (RIGHT | REV_CIGAR | EXTZ_ONLY), on repeats where left and right placement differ."""
    rng = np.random.default_rng(99 + w)
    pairs = []
    for L in (900, 2300, 4000, 6100):
        unit = synth.random_seq(rng, int(rng.integers(2, 9)))
        ref = synth.random_seq(rng, L).copy()
        for _ in range(6):                                   # tandem repeats: a gap can slide inside them
            k = int(rng.integers(0, L - 200)); n = int(rng.integers(20, 120))
            ref[k:k + n] = np.resize(unit, n)
        q = synth.mutate(rng, ref, 0.004, 0.01, 0.01)
        pairs.append((q, ref))
    flags = np.array([_abi.EZ_RIGHT, _abi.EZ_RIGHT | _abi.EZ_REV_CIGAR | _abi.EZ_EXTZ_ONLY, _abi.EZ_RIGHT | _abi.EZ_EXTZ_ONLY,
                      _abi.EZ_RIGHT | _abi.EZ_REV_CIGAR], dtype=np.int32)
    from focalsv_b200.presets import PRESETS
    g = synth._pack("right." + preset, preset, pairs, w, PRESETS[preset].zdrop, flags=flags)
    before = aligner.stats()["exact_path_tasks"]
    bad, ores, gres = compare_group(oracle, aligner, g)
    assert not bad, bad
    assert aligner.stats()["exact_path_tasks"] == before
    # and the placement really differs from the left-aligned run for at least one task
    g2 = g._replace(tasks=g.tasks.copy()); g2.tasks["flag"] &= ~_abi.EZ_RIGHT
    res_l, cig_l = aligner.align_batch(g2.scoring, g2.qarena, g2.tarena, g2.tasks)
    res_r, cig_r = aligner.align_batch(g.scoring, g.qarena, g.tarena, g.tasks)
    assert any(not np.array_equal(task_cigar(res_l[i], cig_l), task_cigar(res_r[i], cig_r)) for i in range(len(res_l)))


@pytest.mark.parametrize("preset,w", [("hifiasm", 500), ("asm5", 3001), ("map-ont", 60)])
def test_approximate_maximum_score_only_on_the_dpx_kernel(oracle, aligner, preset, w):
    """KSW_EZ_SCORE_ONLY | KSW_EZ_APPROX_MAX (| KSW_EZ_APPROX_DROP), hifiasm's call (Correct.cpp:7811,7956;
    ksw2_extz2_sse.c:270-286): one cell is followed greedily, z-drop only with APPROX_DROP, mqe / mte never set."""
    rng = np.random.default_rng(55 + w)
    pairs, flags = [], []
    base = _abi.EZ_SCORE_ONLY | _abi.EZ_APPROX_MAX
    for i, L in enumerate([300, 1200, 2500, 4100, 5000, 6300, 800, 3300]):
        ref = synth.random_seq(rng, L)
        q = synth.mutate(rng, ref, 0.02, 0.01, 0.01)
        if i % 4 == 1:                                   # diverges half way: APPROX_DROP must stop, plain approx must not
            q = np.concatenate([q[:len(q) // 2], synth.random_seq(rng, len(q) // 2)])
        if i % 4 == 2:
            q = np.concatenate([q[:L // 3], synth.random_seq(rng, 90), q[L // 3:]])
        pairs.append((q, ref))
        flags.append(base | (_abi.EZ_APPROX_DROP if i % 2 else 0) | (_abi.EZ_EXTZ_ONLY if i % 3 == 0 else 0))
    from focalsv_b200.presets import PRESETS
    g = synth._pack("approx." + preset, preset, pairs, w, PRESETS[preset].zdrop, flags=np.array(flags, dtype=np.int32))
    before = aligner.stats()["exact_path_tasks"]
    bad, ores, gres = compare_group(oracle, aligner, g)
    assert not bad, bad
    assert aligner.stats()["exact_path_tasks"] == before
    assert int(gres["zdropped"].sum()) >= 1 and all(int(x) == _abi.NEG_INF for x in gres["mqe"])


@pytest.mark.parametrize("preset,w", [("hifiasm", 500), ("asm5", 3001)])
def test_segmented_long_tasks(oracle, preset, w):
    """Long tasks cut into cold-started segments that run on different CTAs (fsv_common.cuh, DevSeg): every field and
    CIGAR word equals the oracle's, whether the boundary check passes (mutated sequences), fails and the upper segment is
    repaired from its predecessor's end state (identical sequences keep a cold-start artefact at EVERY boundary), the
    alignment z-drops inside a later segment, or short tasks share the batch."""
    from focalsv_b200 import api
    from focalsv_b200.presets import PRESETS
    rng = np.random.default_rng(808 + w)
    L = 70000
    pairs, flags = [], []
    ref = synth.random_seq(rng, L)
    q, _ = synth.plant_svs(rng, ref, 5, max_net=min(w // 2 - 50, 1200), max_len=min(w // 2 - 60, 1000))
    pairs.append((synth.mutate(rng, q, 0.001, 0.0003, 0.0003), ref)); flags.append(0)
    ref = synth.random_seq(rng, L + 3000)
    pairs.append((ref.copy(), ref)); flags.append(0)                                          # identical: repair path
    ref = synth.random_seq(rng, L)
    q = synth.mutate(rng, ref, 0.001, 0.0003, 0.0003)
    q = np.concatenate([q[: 2 * L // 3], synth.random_seq(rng, L // 3)])                      # diverges at 2/3: z-drop in a later segment
    pairs.append((q, ref)); flags.append(0)
    ref = synth.random_seq(rng, L)
    pairs.append((synth.mutate(rng, ref, 0.002, 0.001, 0.001), ref)); flags.append(_abi.EZ_EXTZ_ONLY | _abi.EZ_REV_CIGAR)
    for fl in (_abi.EZ_EXTZ_ONLY, 0):                                                         # diverges at 1/4: the segments behind the
        ref = synth.random_seq(rng, L)                                                        # z-drop run on unrelated sequence, where a cold
        q = synth.mutate(rng, ref, 0.001, 0.0003, 0.0003)                                     # start need not converge; they must not be
        q = np.concatenate([q[: L // 4], synth.random_seq(rng, 3 * L // 4)])                  # looked at (no repair)   
        pairs.append((q, ref)); flags.append(fl)
    ref = synth.random_seq(rng, L)                                                            # |qlen - tlen| > w: the band runs out before the
    q = np.concatenate([synth.mutate(rng, ref, 0.001, 0.0003, 0.0003), synth.random_seq(rng, 60000 if w == 500 else 140000)])
    pairs.append((q, ref)); flags.append(0)                                                   # end and whole segments start behind it
    n_long = len(pairs)
    for Ls in (300, 2500, 9000):                                                              # short company
        ref = synth.random_seq(rng, Ls)
        pairs.append((synth.mutate(rng, ref, 0.01, 0.004, 0.004), ref)); flags.append(0)
    g = synth._pack("seg." + preset, preset, pairs, w, PRESETS[preset].zdrop, flags=np.array(flags, dtype=np.int32))
    al = api.Aligner(0)
    try:
        al.set_option("segment_min_diags", 50000)          # force: every long task is segmented, extensions included
        bad, ores, gres = compare_group(oracle, al, g, threads=16)
        assert not bad, bad
        assert int(gres["zdropped"][2]) == 1 and int(gres["zdropped"][4]) == 1 and int(gres["zdropped"][5]) == 1
        st = al.stats()
        assert st["segmented_tasks"] == n_long
        assert st["segment_fallbacks"] <= 2 * 16, st         # repaired segments (the identical pair: one per boundary at most)
        launches_seg = al.stats()["fill_launches"]
        # a cold-start lead of half the band is far too short: most boundaries fail their check and the upper segments are
        # run again from their predecessors' end states (the repair path), with the same results
        al.set_option("segment_warm_pct", 50)
        bad, _, _ = compare_group(oracle, al, g, threads=16)
        assert not bad, bad
        assert al.stats()["segment_fallbacks"] >= (3 if w > 1000 else 0), al.stats()      # (a band of 500 forgets its start within the 1 024 extra rows)
        al.set_option("segment_warm_pct", 300)
        launches_seg = al.stats()["fill_launches"] - launches_seg
        before = al.stats()["fill_launches"]
        al.set_option("segment_min_diags", 0)               # and the same batch unsegmented
        bad, _, _ = compare_group(oracle, al, g, threads=16)
        assert not bad, bad
        assert al.stats()["fill_launches"] - before < launches_seg            # the segmented run had the extra segment launch
    finally:
        al.close()


def test_segmented_tasks_queue_for_pool_pages(oracle):
    """More long tasks than the traceback pool holds at once: a segmented task takes all its pages when its first segment
    starts, in queue order (fsv_fill_dpx.cuh, seg_admit), so the later ones wait on the device for earlier CIGARs to be written."""
    from focalsv_b200 import api
    from focalsv_b200.presets import PRESETS
    rng = np.random.default_rng(4242)
    pairs, flags = [], []
    for i in range(7):
        L = 70000 - 3000 * i
        ref = synth.random_seq(rng, L)
        q, _ = synth.plant_svs(rng, ref, 4, max_net=180, max_len=170)
        pairs.append((synth.mutate(rng, q, 0.001, 0.0003, 0.0003), ref)); flags.append(_abi.EZ_EXTZ_ONLY if i == 3 else 0)
    for Ls in (400, 3000):
        ref = synth.random_seq(rng, Ls)
        pairs.append((synth.mutate(rng, ref, 0.01, 0.004, 0.004), ref)); flags.append(0)
    g = synth._pack("queue", "hifiasm", pairs, 500, PRESETS["hifiasm"].zdrop, flags=np.array(flags, dtype=np.int32))
    al = api.Aligner(0)
    try:
        al.set_option("traceback_budget_bytes", 8 * (32 << 20))       # 8 pages; a long task needs 3: two at a time
        al.set_option("segment_min_diags", 50000)
        for align in (0, 1):
            al.set_option("segment_align_pages", align)
            bad, ores, gres = compare_group(oracle, al, g, threads=16)
            assert not bad, bad
            st = al.stats()
            assert st["segmented_tasks"] == 7, st
            assert st["segment_fallbacks"] == 0, st
    finally:
        al.close()


def test_mixed_kernel_variants_share_the_pool(oracle, aligner):
    """One batch whose tasks land on several kernel variants at once (1/2/4-warp DPX classes, score-only and
    CIGAR, wildcard tasks, and right-aligned tasks on the general kernel), all running concurrently on one page pool."""
    rng = np.random.default_rng(321)
    pairs, flags = [], []
    for i in range(48):
        L = int(rng.choice([300, 900, 2500, 5000]))
        ref = synth.random_seq(rng, L)
        q = synth.mutate(rng, ref, 0.01, 0.004, 0.004)
        if i % 7 == 0:
            q = q.copy(); q[:: 50] = 4          # wildcard bases
        pairs.append((q, ref)); flags.append([0, _abi.EZ_SCORE_ONLY, _abi.EZ_EXTZ_ONLY, _abi.EZ_RIGHT][i % 4])
    g = synth._pack("mixed", "map-ont", pairs, -1, 400, flags=np.array(flags, dtype=np.int32))
    tasks = g.tasks.copy()
    tasks["w"] = np.where(np.arange(len(tasks)) % 3 == 0, 200, np.where(np.arange(len(tasks)) % 3 == 1, 900, 1800))
    g = g._replace(tasks=tasks)
    bad, ores, gres = compare_group(oracle, aligner, g)
    assert not bad, bad[:5]
