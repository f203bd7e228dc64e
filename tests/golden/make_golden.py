"""Generate the golden known-answer vectors under tests/golden/ (run HERE, where /root/reference exists).

single-affine vectors: produced by the reference's OWN ksw_extz2_sse, compiled in place from
  /root/reference/software/hifiasm-0.16.1/ksw2_extz2_sse.c (oracle/_ref/libksw2_ref.so, oracle/build.py);
  the setups follow SURVEY.md appendix C (smoke, KAT1..KAT6x) plus band-edge / wildcard / flag cases.
dual-affine vectors: minimap2's ksw_extd2_sse is not under /root/reference, so these are frozen outputs of
  the oracle's restatement (oracle/ksw2_oracle.c), kept so that any later change of either side shows up.

Usage: python tests/golden/make_golden.py   -> tests/golden/kat_extz2.npz, tests/golden/kat_extd2.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from focalsv_b200 import _abi, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def cases(dual):
    rng = np.random.default_rng(4242 + dual)
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    out = []
    sc_h = (2, 4, 4, 2, -1, -1, 0)            # hifiasm constants (Correct.h:1194-1199), N scores 0
    sc_map = (2, 4, 4, 2, 24, 1, 1)           # map-pb / map-ont
    sc_asm5 = (1, 19, 39, 3, 81, 1, 1)
    sc_asm10 = (1, 9, 16, 2, 41, 1, 1)
    sc_hifi = (1, 4, 6, 2, 26, 1, 1)
    scs = [sc_map, sc_asm5, sc_asm10, sc_hifi] if dual else [sc_h]
    t = np.array([code[c] for c in "ACGTACGTTTGACCAGTAGGACCATTTACGATCAGGGATACCA"], dtype=np.uint8)
    q = np.array([code[c] for c in "ACGTACGTTTGACAGTAGGACCATTTAACGATCAGCGATACCA"], dtype=np.uint8)
    out.append(dict(name="smoke", q=q, t=t, sc=scs[0], w=-1, zdrop=400, end_bonus=0, flag=0))

    def pair(tl, sub, ins, dele, sv=None):
        ref = synth.random_seq(rng, tl)
        qq = synth.mutate(rng, ref, sub, ins, dele)
        if sv == "del":
            qq = np.concatenate([qq[:1000], qq[1300:]])
        if sv == "ins":
            qq = np.concatenate([qq[:1000], synth.random_seq(rng, 300), qq[1000:]])
        if sv == "rand":
            qq = np.concatenate([qq[:1500], synth.random_seq(rng, len(qq) - 1500)])
        return qq, ref

    F = _abi
    for k, sc in enumerate(scs):
        q1, t1 = pair(2000, 0.01, 0, 0, "del")
        out.append(dict(name="kat1_del_%d" % k, q=q1, t=t1, sc=sc, w=500, zdrop=400, end_bonus=0, flag=0))
        out.append(dict(name="kat5_band50_%d" % k, q=q1, t=t1, sc=sc, w=50, zdrop=-1, end_bonus=0, flag=0))
        q2, t2 = pair(2000, 0.01, 0, 0, "ins")
        out.append(dict(name="kat2_ins_%d" % k, q=q2, t=t2, sc=sc, w=500, zdrop=400, end_bonus=0, flag=0))
        q3, t3 = pair(1500, 0.03, 0.01, 0.01)
        out.append(dict(name="kat3_ext_right_rev_%d" % k, q=q3, t=t3, sc=sc, w=500, zdrop=400, end_bonus=10,
                        flag=F.EZ_EXTZ_ONLY | F.EZ_RIGHT | F.EZ_REV_CIGAR))
        q4, t4 = pair(3000, 0.01, 0, 0, "rand")
        out.append(dict(name="kat4_zdrop_%d" % k, q=q4, t=t4, sc=sc, w=500, zdrop=100, end_bonus=0, flag=F.EZ_EXTZ_ONLY))
        q6, t6 = pair(5000, 0.08, 0.04, 0.03)
        out.append(dict(name="kat6_approx_%d" % k, q=q6, t=t6, sc=sc, w=500, zdrop=400, end_bonus=0,
                        flag=F.EZ_SCORE_ONLY | F.EZ_APPROX_MAX | F.EZ_APPROX_DROP))
        out.append(dict(name="kat6x_exact_%d" % k, q=q6, t=t6, sc=sc, w=500, zdrop=400, end_bonus=0, flag=F.EZ_SCORE_ONLY))
        # band edges, wildcards, generic scoring, ragged lengths
        q7, t7 = pair(700, 0.05, 0.02, 0.02)
        for w in (1, 7, 16, 33):
            out.append(dict(name="band%d_%d" % (w, k), q=q7, t=t7, sc=sc, w=w, zdrop=200, end_bonus=0, flag=0))
        qn, tn = q7.copy(), t7.copy()
        qn[::37] = 4
        tn[5::53] = 4
        out.append(dict(name="wildcard_%d" % k, q=qn, t=tn, sc=sc, w=100, zdrop=200, end_bonus=0, flag=0))
        out.append(dict(name="generic_%d" % k, q=qn, t=tn, sc=sc, w=100, zdrop=200, end_bonus=0, flag=F.EZ_GENERIC_SC))
        out.append(dict(name="ragged_%d" % k, q=q7[:97], t=t7[:411], sc=sc, w=-1, zdrop=-1, end_bonus=5, flag=F.EZ_EXTZ_ONLY))
        out.append(dict(name="one_base_%d" % k, q=q7[:1], t=t7[:1], sc=sc, w=-1, zdrop=-1, end_bonus=0, flag=0))
    return out


def main():
    for dual in (False, True):
        cs = cases(dual)
        blob = {"names": np.array([c["name"] for c in cs])}
        res = np.zeros(len(cs), dtype=_abi.RESULT_DTYPE)
        for i, c in enumerate(cs):
            a, b, q, e, q2, e2, amb = c["sc"]
            sc = _abi.make_scoring(a, b, q, e, q2, e2, sc_ambi=amb)
            if dual:
                r, cig = O.extd2(c["q"], c["t"], sc, w=c["w"], zdrop=c["zdrop"], end_bonus=c["end_bonus"], flag=c["flag"])
            else:
                r, cig = O.ref_extz2(c["q"], c["t"], sc, w=c["w"], zdrop=c["zdrop"], end_bonus=c["end_bonus"], flag=c["flag"])
            res[i] = r
            blob["q%d" % i] = c["q"]; blob["t%d" % i] = c["t"]; blob["cigar%d" % i] = cig
            blob["par%d" % i] = np.array(list(c["sc"]) + [c["w"], c["zdrop"], c["end_bonus"], c["flag"]], dtype=np.int32)
        blob["results"] = res
        path = os.path.join(HERE, "kat_extd2.npz" if dual else "kat_extz2.npz")
        np.savez_compressed(path, **blob)
        print(path, len(cs), "cases;", "source:", "oracle restatement" if dual else "reference ksw_extz2_sse compiled in place")


if __name__ == "__main__":
    main()
