"""Helpers shared by the parity tests."""
import numpy as np

from focalsv_b200 import _abi
from focalsv_b200._abi import EZ_FIELDS, cigar_str


def mutate_simple(rng, t, sub, ins, dele):
    from focalsv_b200.synth import mutate
    return mutate(rng, t, sub, ins, dele)


def random_case(rng, max_len=400, dual=False, allow_n=True):
    """One random task: sequences, scoring, band, z-drop, flags (the shapes of tests' fuzzers)."""
    tl = int(rng.integers(1, max_len))
    t = rng.integers(0, 4, tl).astype(np.uint8)
    mode = int(rng.integers(4))
    if mode == 0:
        q = rng.integers(0, 4, int(rng.integers(1, max_len))).astype(np.uint8)
    else:
        q = mutate_simple(rng, t, rng.random() * 0.1, rng.random() * 0.05, rng.random() * 0.05)
        if rng.random() < 0.3 and len(q) > 20:
            k = int(rng.integers(0, len(q) - 10))
            L = int(rng.integers(1, 60))
            if rng.random() < 0.5:
                q = np.concatenate([q[:k], q[k + L:]])
            else:
                q = np.concatenate([q[:k], rng.integers(0, 4, L).astype(np.uint8), q[k:]])
        if len(q) == 0:
            q = np.array([0], dtype=np.uint8)
    if allow_n and rng.random() < 0.1:
        q = q.copy()
        q[rng.integers(0, len(q), max(1, len(q) // 20))] = 4
    if allow_n and rng.random() < 0.1:
        t = t.copy()
        t[rng.integers(0, len(t), max(1, len(t) // 20))] = 4
    a = int(rng.integers(1, 4)); b = int(rng.integers(1, 8))
    gq = int(rng.integers(1, 10)); ge = int(rng.integers(1, 4))
    amb = int(rng.integers(0, 3))
    if dual:
        gq2 = int(gq + rng.integers(0, 30)); ge2 = int(max(1, ge - rng.integers(0, 2)))
        if rng.random() < 0.2:
            gq, gq2, ge, ge2 = gq2, gq, ge2, ge      # exercise the piece swap
        sc = _abi.make_scoring(a, b, gq, ge, gq2, ge2, sc_ambi=amb)
    else:
        sc = _abi.make_scoring(a, b, gq, ge, sc_ambi=amb)
    w = int(rng.choice([-1, 1, 3, 5, 10, 17, 33, 50, 100, 500]))
    zd = int(rng.choice([-1, 10, 50, 100, 400]))
    flag = 0
    for f, pb in ((_abi.EZ_SCORE_ONLY, 0.25), (_abi.EZ_RIGHT, 0.3), (_abi.EZ_GENERIC_SC, 0.15),
                  (_abi.EZ_APPROX_MAX, 0.15), (_abi.EZ_EXTZ_ONLY, 0.4), (_abi.EZ_REV_CIGAR, 0.3)):
        if rng.random() < pb:
            flag |= f
    if flag & _abi.EZ_APPROX_MAX and rng.random() < 0.5:
        flag |= _abi.EZ_APPROX_DROP
    eb = int(rng.choice([0, 0, 5, 10, -1]))
    return dict(q=q, t=t, sc=sc, w=w, zdrop=zd, end_bonus=eb, flag=flag)


def same_result(r1, c1, r2, c2, fields=EZ_FIELDS):
    return all(int(r1[f]) == int(r2[f]) for f in fields) and np.array_equal(np.asarray(c1), np.asarray(c2))


def describe(r, c):
    return "%s %s" % ({f: int(r[f]) for f in EZ_FIELDS}, cigar_str(c)[:120])


def oracle_batch(O, group, threads=8):
    """Run a synth.Group through the oracle; returns (results, list of cigars)."""
    res, arena = O.run_batch(group.scoring, group.qarena, group.tarena, group.tasks, threads=threads)
    cigs = [arena[int(r["cigar_off"]):int(r["cigar_off"]) + int(r["n_cigar"])] for r in res]
    return res, cigs


def compare_group(O, al, group, threads=8):
    """GPU (through the C ABI) vs oracle on one group; returns list of mismatching task indices."""
    from focalsv_b200.api import task_cigar
    ores, ocigs = oracle_batch(O, group, threads)
    gres, garena = al.align_batch(group.scoring, group.qarena, group.tarena, group.tasks)
    bad = []
    for i in range(len(group.tasks)):
        gc = task_cigar(gres[i], garena)
        if not same_result(ores[i], ocigs[i], gres[i], gc) or int(ores[i]["cells"]) != int(gres[i]["cells"]):
            bad.append(i)
    return bad, ores, gres


class OracleRunner(object):
    """align_batch on the CPU oracle: lets host-side code that takes "anything with align_batch" (hook.realign_regions,
    hook.realign_regions_chained, dropin.align_fastas) be exercised without a GPU.  Test infrastructure."""
    def __init__(self, O, threads=8):
        self.O, self.threads = O, threads

    def align_batch(self, sc, qa, ta, tasks):
        return self.O.run_batch(sc, qa, ta, tasks, threads=self.threads)
