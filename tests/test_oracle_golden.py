"""The oracle against the golden known-answer vectors (tests/golden/, made by make_golden.py).

kat_extz2.npz holds outputs of the reference's own ksw_extz2_sse compiled in place; kat_extd2.npz holds
frozen outputs of the dual-affine restatement (the reference's dual-affine body is not in its tree)."""
import os

import numpy as np
import pytest

from focalsv_b200 import _abi
from util import describe, same_result

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLD, name))
    n = len(z["names"])
    cases = []
    for i in range(n):
        par = z["par%d" % i]
        a, b, q, e, q2, e2, amb, w, zd, eb, flag = [int(x) for x in par]
        cases.append(dict(name=str(z["names"][i]), q=z["q%d" % i], t=z["t%d" % i], cigar=z["cigar%d" % i],
                          sc=_abi.make_scoring(a, b, q, e, q2, e2, sc_ambi=amb), w=w, zdrop=zd, end_bonus=eb, flag=flag,
                          res=z["results"][i]))
    return cases


@pytest.mark.parametrize("fname,dual", [("kat_extz2.npz", False), ("kat_extd2.npz", True)])
def test_oracle_matches_golden(oracle, fname, dual):
    for c in load(fname):
        f = oracle.extd2 if dual else oracle.extz2
        r, cig = f(c["q"], c["t"], c["sc"], w=c["w"], zdrop=c["zdrop"], end_bonus=c["end_bonus"], flag=c["flag"])
        assert same_result(c["res"], c["cigar"], r, cig), (c["name"], describe(c["res"], c["cigar"]), describe(r, cig))


def test_survey_appendix_c_smoke(oracle):
    """SURVEY.md appendix C, first row: score 66, max_q = max_t = 42, cigar 12M1D14M1I16M."""
    c = load("kat_extz2.npz")[0]
    assert c["name"] == "smoke"
    r, cig = oracle.extz2(c["q"], c["t"], c["sc"], w=-1, zdrop=400)
    assert int(r["score"]) == 66 and int(r["max"]) == 66 and int(r["max_q"]) == 42 and int(r["max_t"]) == 42
    assert _abi.cigar_str(cig) == "12M1D14M1I16M"
    assert int(r["mte_q"]) == 37        # r - rounded en: the reference's 16-lane quirk (ksw2_extz2_sse.c:263-264)
