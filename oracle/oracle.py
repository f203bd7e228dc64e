"""ctypes wrapper around the CPU oracle.  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs — never by focalsv_b200 (the product).
"""
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from focalsv_b200._abi import RESULT_DTYPE, TASK_DTYPE, Scoring, scoring_mat  # noqa: E402

_u8p = C.POINTER(C.c_uint8)
_i8p = C.POINTER(C.c_int8)
_u32p = C.POINTER(C.c_uint32)


class Diag(C.Structure):
    _fields_ = [("clamp_inband", C.c_int64), ("clamp_oob", C.c_int64), ("clamp_top", C.c_int64),
                ("wraps", C.c_int64)]


def _paths():
    port = os.path.join(_HERE, "_ref", "libfsv_oracle.so")
    ref = os.path.join(_HERE, "_ref", "libksw2_ref.so")
    need_build = not os.path.exists(port) or (os.path.isdir("/root/reference") and not os.path.exists(ref))
    if not need_build:
        src = os.path.join(_HERE, "ksw2_oracle.c")
        need_build = os.path.getmtime(src) > os.path.getmtime(port)
    if need_build:
        sys.path.insert(0, _HERE)
        import build as _b
        _b.build()
    return port, (ref if os.path.exists(ref) else None)


_PORT, _REF = _paths()
port = C.CDLL(_PORT)
ref = C.CDLL(_REF) if _REF else None

port.fsvo_task_cells.restype = C.c_int64
port.fsvo_gotoh2_global.restype = C.c_int32
port.fsvo_edit_distance.restype = C.c_int32
port.fsvo_score_cigar.restype = C.c_int32


def have_reference():
    return ref is not None


def _ptr(a, typ):
    return a.ctypes.data_as(typ)


def _as_u8(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint8))


def _run1(fn, dual, query, target, sc, w, zdrop, end_bonus, flag, want_diag):
    query, target = _as_u8(query), _as_u8(target)
    mat = scoring_mat(sc)
    out = np.zeros(1, dtype=RESULT_DTYPE)
    cap = len(query) + len(target) + 4
    cig = np.zeros(cap, dtype=np.uint32)
    dg = Diag()
    args = [C.c_int(len(query)), _ptr(query, _u8p), C.c_int(len(target)), _ptr(target, _u8p),
            C.c_int8(sc.m), _ptr(mat, _i8p), C.c_int8(sc.q), C.c_int8(sc.e)]
    if dual:
        args += [C.c_int8(sc.q2), C.c_int8(sc.e2)]
    args += [C.c_int(w), C.c_int(zdrop), C.c_int(end_bonus), C.c_int(flag), out.ctypes.data_as(C.c_void_p),
             _ptr(cig, _u32p), C.c_int(cap)]
    if want_diag is not None:
        args.append(C.byref(dg) if want_diag else None)
    n = fn(*args)
    if n < 0:
        raise MemoryError("oracle allocation failed")
    return out[0], cig[:n].copy(), dg


def extz2(query, target, sc, w=-1, zdrop=-1, end_bonus=0, flag=0, diag=False):
    """Plain-C restatement of ksw_extz2_sse (ksw2_extz2_sse.c:23-304)."""
    r, c, d = _run1(port.fsvo_extz2, False, query, target, sc, w, zdrop, end_bonus, flag, diag)
    return (r, c, d) if diag else (r, c)


def extd2(query, target, sc, w=-1, zdrop=-1, end_bonus=0, flag=0, diag=False):
    """Plain-C restatement of ksw_extd2_sse (prototype ksw2.h:60-61)."""
    r, c, d = _run1(port.fsvo_extd2, True, query, target, sc, w, zdrop, end_bonus, flag, diag)
    return (r, c, d) if diag else (r, c)


def align(query, target, sc, **kw):
    return extd2(query, target, sc, **kw) if sc.q2 >= 0 else extz2(query, target, sc, **kw)


def ref_extz2(query, target, sc, w=-1, zdrop=-1, end_bonus=0, flag=0):
    """The reference's own compiled ksw_extz2_sse (oracle/_ref/libksw2_ref.so)."""
    if ref is None:
        raise RuntimeError("oracle/_ref/libksw2_ref.so not built (needs /root/reference)")
    r, c, _ = _run1(ref.fsvref_extz2, False, query, target, sc, w, zdrop, end_bonus, flag, None)
    return r, c


def task_cells(qlen, tlen, w):
    return int(port.fsvo_task_cells(C.c_int(qlen), C.c_int(tlen), C.c_int(w)))


def gotoh2_global(query, target, sc):
    query, target = _as_u8(query), _as_u8(target)
    mat = scoring_mat(sc)
    q2, e2 = (sc.q2, sc.e2) if sc.q2 >= 0 else (sc.q, sc.e)
    return int(port.fsvo_gotoh2_global(C.c_int(len(query)), _ptr(query, _u8p), C.c_int(len(target)),
                                       _ptr(target, _u8p), C.c_int(sc.m), _ptr(mat, _i8p),
                                       C.c_int(sc.q), C.c_int(sc.e), C.c_int(q2), C.c_int(e2)))


def edit_distance(a, b):
    """Global unit-cost edit distance of two byte sequences (edlib.align(a, b)["editDistance"], mode NW)."""
    a = np.frombuffer(a.encode() if isinstance(a, str) else bytes(a), dtype=np.uint8) if not isinstance(a, np.ndarray) else _as_u8(a)
    b = np.frombuffer(b.encode() if isinstance(b, str) else bytes(b), dtype=np.uint8) if not isinstance(b, np.ndarray) else _as_u8(b)
    return int(port.fsvo_edit_distance(C.c_int(len(a)), _ptr(a, _u8p), C.c_int(len(b)), _ptr(b, _u8p)))


def score_cigar(query, target, sc, cigar):
    query, target = _as_u8(query), _as_u8(target)
    cigar = np.ascontiguousarray(cigar, dtype=np.uint32)
    mat = scoring_mat(sc)
    qu, tu = C.c_int(0), C.c_int(0)
    s = port.fsvo_score_cigar(C.c_int(len(query)), _ptr(query, _u8p), C.c_int(len(target)), _ptr(target, _u8p),
                              C.c_int(sc.m), _ptr(mat, _i8p), C.c_int(sc.q), C.c_int(sc.e), C.c_int(sc.q2),
                              C.c_int(sc.e2), _ptr(cigar, _u32p), C.c_int(len(cigar)), C.byref(qu), C.byref(tu))
    return int(s), qu.value, tu.value


def run_batch(sc, qarena, tarena, tasks, threads=1, use_reference=False, cigar_cap=None):
    """Thread-pool batch run (one task per thread at a time).  Returns (results, cigar_arena)."""
    qarena, tarena = _as_u8(qarena), _as_u8(tarena)
    tasks = np.ascontiguousarray(tasks, dtype=TASK_DTYPE)
    n = len(tasks)
    out = np.zeros(n, dtype=RESULT_DTYPE)
    if cigar_cap is None:
        cigar_cap = int((tasks["qlen"].astype(np.int64) + tasks["tlen"]).sum()) + 16
    cig = np.zeros(cigar_cap, dtype=np.uint32)
    used = C.c_int64(0)
    if use_reference:
        if ref is None or sc.q2 >= 0:
            raise RuntimeError("compiled reference covers the single-affine kernel only")
        fn = ref.fsvref_run_batch
    else:
        fn = port.fsvo_run_batch
    fn(C.byref(sc), _ptr(qarena, _u8p), _ptr(tarena, _u8p), tasks.ctypes.data_as(C.c_void_p), C.c_int64(n),
       C.c_int(threads), out.ctypes.data_as(C.c_void_p), _ptr(cig, _u32p), C.c_int64(cigar_cap), C.byref(used))
    return out, cig[:used.value].copy()
