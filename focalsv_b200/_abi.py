"""ctypes / numpy mirrors of include/focalsv_cuda.h (ABI version 4).

Kept in one place so that the product binding (focalsv_b200.api), the oracle
wrapper (oracle/oracle.py) and the tests all see the same layouts.
"""
import ctypes as C

import numpy as np

ABI_VERSION = 4
NEG_INF = -0x40000000

# ksw2.h:8-14
EZ_SCORE_ONLY = 0x01
EZ_RIGHT = 0x02
EZ_GENERIC_SC = 0x04
EZ_APPROX_MAX = 0x08
EZ_APPROX_DROP = 0x10
EZ_EXTZ_ONLY = 0x40
EZ_REV_CIGAR = 0x80

PLAN_DPX, PLAN_GENERAL, PLAN_SEGMENTED, PLAN_EXCLUSIVE, PLAN_EDGE_WARP = 1, 2, 4, 8, 16

OK = 0
ERR_NO_DEVICE = -1
ERR_CUDA = -2
ERR_INVALID = -3
ERR_NOMEM = -4
ERR_CIGAR_CAP = -5
ERR_SCORING = -6
ERR_STATE = -7


class Scoring(C.Structure):
    _fields_ = [("m", C.c_int8), ("q", C.c_int8), ("e", C.c_int8), ("q2", C.c_int8), ("e2", C.c_int8),
                ("reserved", C.c_int8 * 3), ("mat", C.c_int8 * 32)]


TASK_DTYPE = np.dtype([("q_off", "<i8"), ("t_off", "<i8"), ("qlen", "<i4"), ("tlen", "<i4"), ("w", "<i4"),
                       ("zdrop", "<i4"), ("end_bonus", "<i4"), ("flag", "<i4")], align=True)
RESULT_DTYPE = np.dtype([("max", "<i4"), ("zdropped", "<i4"), ("max_q", "<i4"), ("max_t", "<i4"),
                         ("mqe", "<i4"), ("mqe_t", "<i4"), ("mte", "<i4"), ("mte_q", "<i4"),
                         ("score", "<i4"), ("reach_end", "<i4"), ("n_cigar", "<i4"), ("status", "<i4"),
                         ("cigar_off", "<i8"), ("cells", "<i8")], align=True)
SIGNATURE_DTYPE = np.dtype([("task", "<i4"), ("svtype", "<i4"), ("pos", "<i8"), ("svlen", "<i4"), ("read_start", "<i4"),
                            ("read_end", "<i4"), ("pad_", "<i4")], align=True)
PAIR_DTYPE = np.dtype([("a_off", "<i8"), ("b_off", "<i8"), ("a_len", "<i4"), ("b_len", "<i4")], align=True)
RECORD_DTYPE = np.dtype([("pos", "<i8"), ("ref_end", "<i8"), ("cigar_off", "<i8"), ("n_cigar", "<i4"), ("query_length", "<i4"),
                         ("score", "<i4"), ("zdropped", "<i4"), ("is_reverse", "<i4"), ("mapq", "<i4")], align=True)
PIECE_DTYPE = np.dtype([("q_beg", "<i4"), ("q_end", "<i4"), ("t_beg", "<i4"), ("t_end", "<i4")], align=True)
CHAIN_DTYPE = np.dtype([("strand", "<i4"), ("score", "<i4"), ("n_anchors", "<i4"), ("sub_score", "<i4"), ("q_beg", "<i4"), ("q_end", "<i4"),
                        ("t_beg", "<i4"), ("t_end", "<i4"), ("piece_off", "<i4"), ("n_pieces", "<i4"), ("lq", "<i4"), ("lt", "<i4"),
                        ("rq", "<i4"), ("rt", "<i4")], align=True)
assert RECORD_DTYPE.itemsize == 48 and PIECE_DTYPE.itemsize == 16 and CHAIN_DTYPE.itemsize == 56
assert TASK_DTYPE.itemsize == 40 and RESULT_DTYPE.itemsize == 64 and SIGNATURE_DTYPE.itemsize == 32 and PAIR_DTYPE.itemsize == 24

# fields that must be bit-identical to ksw_extz_t (ksw2.h:23-32)
EZ_FIELDS = ("max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "reach_end", "n_cigar")


class PresetC(C.Structure):
    _fields_ = [("name", C.c_char * 16)] + [(k, C.c_int32) for k in
                ("a", "b", "q", "e", "q2", "e2", "zdrop", "zdrop_inv", "bw", "bw_long", "sc_ambi", "end_bonus")]


class ChainOpts(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("k", "w", "max_occ", "max_gap", "min_fill", "max_chains", "min_chain_score", "min_anchors",
                                         "a", "q", "e", "end_bonus")]


class Stats(C.Structure):
    _fields_ = [("tasks", C.c_int64), ("cells", C.c_int64), ("fill_launches", C.c_int64),
                ("backtrack_launches", C.c_int64), ("other_launches", C.c_int64),
                ("exact_path_tasks", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("fill_ms", C.c_double), ("backtrack_ms", C.c_double), ("total_ms", C.c_double),
                ("traceback_bytes", C.c_int64), ("segmented_tasks", C.c_int64), ("segment_fallbacks", C.c_int64)]


def simple_mat(a, b, sc_ambi=1, m=5):
    """5x5 matrix as minimap2's ksw_gen_simple_mat builds it: `a` on the diagonal,
    -b elsewhere, -sc_ambi in the last row/column.  hifiasm's
    call site uses 0 in the wildcard row/column (Correct.cpp:7670-7672)."""
    a, b = abs(int(a)), abs(int(b))
    mat = np.full((m, m), -b, dtype=np.int8)
    for i in range(m - 1):
        mat[i, i] = a
    mat[m - 1, :] = -abs(int(sc_ambi))
    mat[:, m - 1] = -abs(int(sc_ambi))
    return mat.reshape(-1)


def make_scoring(a, b, q, e, q2=-1, e2=-1, sc_ambi=1, m=5, mat=None):
    sc = Scoring()
    sc.m = m
    sc.q, sc.e, sc.q2, sc.e2 = q, e, q2, e2
    mm = simple_mat(a, b, sc_ambi, m) if mat is None else np.asarray(mat, dtype=np.int8).reshape(-1)
    for i in range(m * m):
        sc.mat[i] = int(mm[i])
    return sc


def scoring_mat(sc):
    return np.array([sc.mat[i] for i in range(sc.m * sc.m)], dtype=np.int8)


def cigar_str(words):
    return "".join("%d%s" % (int(w) >> 4, "MIDNSHP=XB"[int(w) & 0xF]) for w in words)
