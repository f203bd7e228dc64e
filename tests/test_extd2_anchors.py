"""Acceptance anchors of the dual-affine restatement, since the reference's
ksw_extd2_sse body is not in its tree ("parity unpinned" at the source level):
  (i)   extd2(q,e,q,e) == extz2(q,e) on every ksw_extz_t field and CIGAR (full band: no stale-lane effects)
  (ii)  the global score equals an independent int32 two-piece Gotoh DP
  (iii) every CIGAR re-scores to ez.score under min(q + l*e, q2 + l*e2)."""
import numpy as np

from focalsv_b200 import _abi, synth
from util import describe, same_result


def _pair(rng, n, err):
    t = synth.random_seq(rng, n)
    q = synth.mutate(rng, t, err * 0.5, err * 0.25, err * 0.25)
    if len(q) > 60 and rng.random() < 0.7:
        k = int(rng.integers(10, len(q) - 40)); L = int(rng.integers(5, 40))
        q = np.concatenate([q[:k], q[k + L:]]) if rng.random() < 0.5 else np.concatenate([q[:k], synth.random_seq(rng, L), q[k:]])
    return q, t


def test_equal_pieces_reduce_to_single_affine(oracle):
    rng = np.random.default_rng(77)
    for it in range(300):
        q, t = _pair(rng, int(rng.integers(5, 300)), rng.random() * 0.15)
        a, b, go, ge = int(rng.integers(1, 4)), int(rng.integers(1, 7)), int(rng.integers(1, 9)), int(rng.integers(1, 4))
        s1 = _abi.make_scoring(a, b, go, ge, sc_ambi=1)
        s2 = _abi.make_scoring(a, b, go, ge, go, ge, sc_ambi=1)
        flag = int(rng.choice([0, _abi.EZ_EXTZ_ONLY, _abi.EZ_RIGHT, _abi.EZ_SCORE_ONLY, _abi.EZ_EXTZ_ONLY | _abi.EZ_REV_CIGAR]))
        zd = int(rng.choice([-1, 50, 400]))
        r1, c1 = oracle.extz2(q, t, s1, w=-1, zdrop=zd, flag=flag)
        r2, c2 = oracle.extd2(q, t, s2, w=-1, zdrop=zd, flag=flag)
        assert same_result(r1, c1, r2, c2), (it, describe(r1, c1), describe(r2, c2))


def test_global_score_equals_independent_gotoh(oracle):
    rng = np.random.default_rng(78)
    presets = [(1, 19, 39, 3, 81, 1), (1, 9, 16, 2, 41, 1), (1, 4, 6, 2, 26, 1), (2, 4, 4, 2, 24, 1)]
    for it in range(120):
        q, t = _pair(rng, int(rng.integers(20, 260)), rng.random() * 0.1)
        a, b, go, ge, go2, ge2 = presets[it % 4]
        sc = _abi.make_scoring(a, b, go, ge, go2, ge2, sc_ambi=1)
        r, cig = oracle.extd2(q, t, sc, w=-1, zdrop=-1, flag=0)
        assert int(r["score"]) == oracle.gotoh2_global(q, t, sc), (it, describe(r, cig))
        s, qu, tu = oracle.score_cigar(q, t, sc, cig)
        assert (s, qu, tu) == (int(r["score"]), len(q), len(t)), (it, s, describe(r, cig))


def test_single_affine_score_equals_gotoh_and_cigar_rescoring(oracle):
    rng = np.random.default_rng(79)
    sc = _abi.make_scoring(2, 4, 4, 2, sc_ambi=1)
    for it in range(80):
        q, t = _pair(rng, int(rng.integers(20, 260)), rng.random() * 0.1)
        r, cig = oracle.extz2(q, t, sc, w=-1, zdrop=-1, flag=0)
        assert int(r["score"]) == oracle.gotoh2_global(q, t, sc)
        s, qu, tu = oracle.score_cigar(q, t, sc, cig)
        assert (s, qu, tu) == (int(r["score"]), len(q), len(t))
