"""The vectorisable row kernel of the dual-affine oracle must equal its scalar lane loop (same file)."""
import ctypes as C

import numpy as np

from util import describe, random_case, same_result


def test_vector_rows_equal_scalar_lanes(oracle):
    flag = C.c_int.in_dll(oracle.port, "fsvo_force_scalar")
    rng = np.random.default_rng(99)
    for it in range(600):
        c = random_case(rng, dual=True)
        kw = dict(w=c["w"], zdrop=c["zdrop"], end_bonus=c["end_bonus"], flag=c["flag"])
        flag.value = 0
        r1, c1 = oracle.extd2(c["q"], c["t"], c["sc"], **kw)
        flag.value = 1
        try:
            r2, c2 = oracle.extd2(c["q"], c["t"], c["sc"], **kw)
        finally:
            flag.value = 0
        assert same_result(r1, c1, r2, c2), (it, describe(r1, c1), describe(r2, c2))
