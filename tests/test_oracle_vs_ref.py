"""Differential test: the plain-C restatement vs the reference's own compiled ksw_extz2_sse
(oracle/_ref/libksw2_ref.so, built in place from /root/reference when it exists)."""
import numpy as np
import pytest

from util import describe, random_case, same_result


def test_extz2_restatement_equals_compiled_reference(oracle):
    if not oracle.have_reference():
        pytest.skip("oracle/_ref/libksw2_ref.so not present (needs /root/reference to build)")
    rng = np.random.default_rng(2024)
    for it in range(1500):
        c = random_case(rng, dual=False)
        r1, c1 = oracle.extz2(c["q"], c["t"], c["sc"], w=c["w"], zdrop=c["zdrop"], end_bonus=c["end_bonus"], flag=c["flag"])
        r2, c2 = oracle.ref_extz2(c["q"], c["t"], c["sc"], w=c["w"], zdrop=c["zdrop"], end_bonus=c["end_bonus"], flag=c["flag"])
        assert same_result(r1, c1, r2, c2), (it, describe(r1, c1), describe(r2, c2))


def test_cell_count_matches_band_definition(oracle):
    for ql, tl, w in ((100, 100, 10), (1, 50, -1), (777, 333, 50), (2000, 2100, 501), (50, 50, 0)):
        n = 0
        ww = max(ql, tl) if w < 0 else w
        for r in range(ql + tl - 1):
            st = max(0, r - ql + 1, (r - ww + 1) >> 1)
            en = min(tl - 1, r, (r + ww) >> 1)
            if st > en:
                break
            n += en - st + 1
        assert oracle.task_cells(ql, tl, w) == n
