"""The edge-warp fill kernel (focalsv_b200/csrc/fsv_fill_ew.cuh) against the CPU oracle, bit-exact, on the shapes that stress what is new in
it: vectors changing form between the edge warp and the main threads (every 32 antidiagonals), bands that grow from / shrink to a single
vector, rectangular tasks whose band runs out (ksw2_extz2_sse.c:111-114), extension tasks that z-drop, scorings whose carries go negative
(single-affine, :146-147), bands of every main-warp class.  The same batches are run with the kernel switched off (fsv_fill_dpx_kernel) too."""
import numpy as np
import pytest

from focalsv_b200 import _abi, api, synth
from util import compare_group

pytestmark = pytest.mark.gpu


def _random_group(rng, n, dual, lens, bands, wide_scores=False):
    pairs, ws, zds, flags, ebs = [], [], [], [], []
    for i in range(n):
        L = int(rng.integers(lens[0], lens[1]))
        ref = synth.random_seq(rng, L)
        kind = int(rng.integers(5))
        if kind == 0:                                   # unrelated sequences of unequal length: the band runs out or the task z-drops
            q = synth.random_seq(rng, int(rng.integers(64, lens[1])))
        elif kind == 1:                                 # a prefix / a much longer query: rectangular
            q = synth.mutate(rng, ref[:max(64, L // int(rng.integers(2, 5)))], 0.01, 0.004, 0.004)
        else:
            q, _ = synth.plant_svs(rng, ref, int(rng.integers(0, 4)), max_net=200, max_len=150)
            q = synth.mutate(rng, q, rng.random() * 0.05, rng.random() * 0.02, rng.random() * 0.02)
        if len(q) < 64:
            q = np.concatenate([q, synth.random_seq(rng, 64)])
        pairs.append((q, ref))
        ws.append(int(rng.choice(bands)))
        zds.append(int(rng.choice([-1, 50, 200, 400])))
        f = 0
        if rng.random() < 0.4:
            f |= _abi.EZ_EXTZ_ONLY
        if rng.random() < 0.2:
            f |= _abi.EZ_REV_CIGAR
        flags.append(f)
        ebs.append(int(rng.choice([0, 0, 10])))
    g = synth._pack("ew", "asm5" if dual else "hifiasm", pairs, 0, 0, flags=np.array(flags, dtype=np.int32))
    g.tasks["w"], g.tasks["zdrop"], g.tasks["end_bonus"] = ws, zds, ebs
    a = int(rng.integers(1, 4)); b = int(rng.integers(1, 9))
    gq = int(rng.integers(1, 40 if wide_scores else 10)); ge = int(rng.integers(1, 6 if wide_scores else 4))
    if dual:
        sc = _abi.make_scoring(a, b, gq, ge, int(gq + rng.integers(0, 40)), int(max(1, ge - rng.integers(0, 2))))
    else:
        sc = _abi.make_scoring(a, b, gq, ge)
    return g._replace(scoring=sc) if hasattr(g, "_replace") else _with_scoring(g, sc)


def _with_scoring(g, sc):
    g.scoring = sc
    return g


def _check(oracle, al, g, expect_edge_warp=True):
    b = al.batch(g.scoring, g.qarena, g.tarena, g.tasks)
    plan = b.plan()
    b.close()
    if expect_edge_warp:
        assert (plan & _abi.PLAN_EDGE_WARP).any(), "no task of the batch was planned for the edge-warp kernel"
    bad, ores, gres = compare_group(oracle, al, g)
    assert not bad, [(i, int(g.tasks["qlen"][i]), int(g.tasks["tlen"][i]), int(g.tasks["w"][i]), int(g.tasks["zdrop"][i]), hex(int(g.tasks["flag"][i]))) for i in bad[:4]]
    al.set_option("ew_kernel", 0)
    try:
        bad2, _, _ = compare_group(oracle, al, g)
    finally:
        al.set_option("ew_kernel", 1)
    assert not bad2
    return plan


@pytest.mark.parametrize("dual", [False, True])
@pytest.mark.parametrize("seed", [1, 2])
def test_fuzz_eligible_shapes(oracle, aligner, dual, seed):
    rng = np.random.default_rng(100 * seed + dual)
    g = _random_group(rng, 60, dual, (64, 2500), [32, 33, 47, 48, 64, 100, 250, 500, 751], wide_scores=(seed == 2))
    _check(oracle, aligner, g)


@pytest.mark.parametrize("dual,w,nwm", [(False, 500, 1), (True, 751, 2), (True, 1500, 4)])
def test_every_main_warp_class(oracle, aligner, dual, w, nwm):
    """Tasks long enough for the band to reach its full width: 1, 2 and 4 main warps (a band-3001 task needs 6: see test_wide_bands_opt_in)."""
    rng = np.random.default_rng(7 + w)
    pairs = []
    for L in (2 * w + 700, 3 * w + 100, 5 * w):
        ref = synth.random_seq(rng, L)
        q, _ = synth.plant_svs(rng, ref, 3, max_net=min(w // 2 - 50, 600), max_len=min(w // 2 - 60, 500))
        pairs.append((synth.mutate(rng, q, 0.004, 0.002, 0.002), ref))
    g = synth._pack("ewc", "asm5" if dual else "hifiasm", pairs, w, 400, flags=np.array([0, _abi.EZ_EXTZ_ONLY, 0], dtype=np.int32))
    plan = _check(oracle, aligner, g)
    assert ((plan & _abi.PLAN_EDGE_WARP) != 0).all()


def test_wide_bands_opt_in(oracle, aligner):
    """ew_kernel = 2 sends band-3001 tasks (6 main warps + the edge warp) there too; by default they stay with fsv_fill_dpx_kernel."""
    rng = np.random.default_rng(5)
    ref = synth.random_seq(rng, 9000)
    q, _ = synth.plant_svs(rng, ref, 3, max_net=1000, max_len=800)
    g = synth._pack("eww", "asm5", [(synth.mutate(rng, q, 0.004, 0.002, 0.002), ref)], 3001, 200)
    b = aligner.batch(g.scoring, g.qarena, g.tarena, g.tasks)
    assert not (b.plan() & _abi.PLAN_EDGE_WARP).any()
    b.close()
    aligner.set_option("ew_kernel", 2)
    try:
        b = aligner.batch(g.scoring, g.qarena, g.tarena, g.tasks)
        assert (b.plan() & _abi.PLAN_EDGE_WARP).all()
        b.close()
        bad, _, _ = compare_group(oracle, aligner, g)
    finally:
        aligner.set_option("ew_kernel", 1)
    assert not bad


def test_identical_and_tiny_eligible_tasks(oracle, aligner):
    """The smallest tasks the kernel takes (64 bases, band 32), identical sequences (every antidiagonal a new maximum), homopolymers (ties everywhere)."""
    rng = np.random.default_rng(9)
    s64 = synth.random_seq(rng, 64)
    s500 = synth.random_seq(rng, 500)
    homo = np.zeros(300, dtype=np.uint8)
    pairs = [(s64, s64), (s500, s500), (homo, homo), (homo[:200], homo), (s500[:250], s500), (s500, s500[100:])]
    for dual in (False, True):
        g = synth._pack("ewt", "asm5" if dual else "hifiasm", pairs, 32, 100, flags=np.array([0, 0, 0, _abi.EZ_EXTZ_ONLY, _abi.EZ_EXTZ_ONLY, 0], dtype=np.int32))
        _check(oracle, aligner, g)
        g = synth._pack("ewt", "asm5" if dual else "hifiasm", pairs, 100, -1)
        _check(oracle, aligner, g)


def test_exclusive_launch_and_lazy_pages(oracle):
    """The kernel's other launch forms: one CTA per SM (force_excl), and tasks whose traceback pages are taken as the task advances
    (small pages, so that a 6 kb task already spans more than lazy_min_pages of them) out of a pool that only fits a few tasks."""
    rng = np.random.default_rng(21)
    pairs = []
    for L in (900, 2500, 6000, 6100):
        ref = synth.random_seq(rng, L)
        q, _ = synth.plant_svs(rng, ref, 2, max_net=150, max_len=120)
        pairs.append((synth.mutate(rng, q, 0.01, 0.004, 0.004), ref))
    al = api.Aligner(0)
    try:
        for preset, w in (("hifiasm", 500), ("map-hifi", 751)):
            from focalsv_b200.presets import PRESETS
            g = synth._pack("ewx", preset, pairs, w, PRESETS[preset].zdrop)
            al.set_option("force_excl", 1)
            b = al.batch(g.scoring, g.qarena, g.tarena, g.tasks)
            plan = b.plan(); b.close()
            assert ((plan & _abi.PLAN_EDGE_WARP) != 0).all() and ((plan & _abi.PLAN_EXCLUSIVE) != 0).all()
            bad, _, _ = compare_group(oracle, al, g)
            assert not bad, ("exclusive", preset, bad)
            al.set_option("force_excl", 0)
        al.set_option("traceback_page_bytes", 65536)
        al.set_option("traceback_budget_bytes", 400 * 65536)
        al.set_option("pool_stall_ms", 5000)
        for lazy in (16, 3):
            al.set_option("lazy_min_pages", lazy)
            g = synth._pack("ewl", "hifiasm", pairs * 6, 500, 400)
            b = al.batch(g.scoring, g.qarena, g.tarena, g.tasks)
            assert ((b.plan() & _abi.PLAN_EDGE_WARP) != 0).any()
            b.close()
            bad, _, _ = compare_group(oracle, al, g)
            assert not bad, ("lazy", lazy, bad)
    finally:
        al.close()
