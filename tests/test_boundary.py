"""The ksw2-shaped C entry points under test.

fsv_ksw_extz2 / fsv_ksw_extd2 take ksw2.h:54-61's arguments one for one; the one native caller in the reference is
afine_gap_alignment (software/hifiasm-0.16.1/Correct.cpp:7658-7705).  GPU tests: (1) the two entry points called
directly through ctypes, argument for argument, against the reference's own compiled ksw_extz2_sse (oracle/_ref) and the
oracle's dual-affine restatement; (2) a C program (tests/c_probe/afine_gap_probe.c, gcc, linked against the shared
object) that reproduces afine_gap_alignment's call sequence and field reads, compared line by line with the same calls on
the compiled reference.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from focalsv_b200 import _abi, api
from util import random_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _call_single(al, dual, q, t, sc, w, zdrop, end_bonus, flag):
    lib = al._lib
    q = np.ascontiguousarray(q, dtype=np.uint8); t = np.ascontiguousarray(t, dtype=np.uint8)
    mat = _abi.scoring_mat(sc)
    ez = np.zeros(1, dtype=_abi.RESULT_DTYPE)
    cap = len(q) + len(t) + 4
    cig = np.zeros(cap, dtype=np.uint32)
    mp = mat.ctypes.data_as(C.POINTER(C.c_int8))
    if dual:
        rc = lib.fsv_ksw_extd2(al._h, len(q), q.ctypes.data, len(t), t.ctypes.data, sc.m, mp, sc.q, sc.e, sc.q2, sc.e2,
                               w, zdrop, end_bonus, flag, ez.ctypes.data, cig.ctypes.data, cap)
    else:
        rc = lib.fsv_ksw_extz2(al._h, len(q), q.ctypes.data, len(t), t.ctypes.data, sc.m, mp, sc.q, sc.e,
                               w, zdrop, end_bonus, flag, ez.ctypes.data, cig.ctypes.data, cap)
    return rc, ez[0], cig[:max(int(ez[0]["n_cigar"]), 0)]


def test_argtypes_declared():
    """host only: the binding declares both ksw2-shaped entry points (they used to be exported but never bound)."""
    lib = api.load_library()
    assert lib.fsv_ksw_extz2.argtypes is not None and len(lib.fsv_ksw_extz2.argtypes) == 16
    assert lib.fsv_ksw_extd2.argtypes is not None and len(lib.fsv_ksw_extd2.argtypes) == 18


@pytest.mark.gpu
def test_fsv_ksw_extz2_vs_compiled_reference(aligner, oracle):
    """ksw2.h:54-55 argument for argument; the checker is the reference's own ksw2_extz2_sse.c compiled in place."""
    if not oracle.have_reference():
        pytest.skip("oracle/_ref/libksw2_ref.so not present")
    rng = np.random.default_rng(20262)
    bad = []
    for it in range(120):
        c = random_case(rng, max_len=300, dual=False)
        rr, rc_ = oracle.ref_extz2(c["q"], c["t"], c["sc"], c["w"], c["zdrop"], c["end_bonus"], c["flag"])
        rc, ez, cig = _call_single(aligner, False, c["q"], c["t"], c["sc"], c["w"], c["zdrop"], c["end_bonus"], c["flag"])
        assert rc == 0
        if not (all(int(ez[f]) == int(rr[f]) for f in _abi.EZ_FIELDS) and np.array_equal(cig, rc_)):
            bad.append(it)
    assert not bad, "fsv_ksw_extz2 differs from the compiled reference on cases %s" % bad[:10]


@pytest.mark.gpu
def test_fsv_ksw_extd2_vs_oracle(aligner, oracle):
    """ksw2.h:60-61 argument for argument (dual-affine: the checker is the oracle's restatement, source-unpinned)."""
    rng = np.random.default_rng(20263)
    bad = []
    for it in range(120):
        c = random_case(rng, max_len=300, dual=True)
        orr, oc = oracle.extd2(c["q"], c["t"], c["sc"], c["w"], c["zdrop"], c["end_bonus"], c["flag"])
        rc, ez, cig = _call_single(aligner, True, c["q"], c["t"], c["sc"], c["w"], c["zdrop"], c["end_bonus"], c["flag"])
        assert rc == 0
        if not (all(int(ez[f]) == int(orr[f]) for f in _abi.EZ_FIELDS) and np.array_equal(cig, oc)):
            bad.append(it)
    assert not bad, "fsv_ksw_extd2 differs from the oracle on cases %s" % bad[:10]


@pytest.mark.gpu
def test_fsv_ksw_argument_errors(aligner):
    q = np.zeros(8, np.uint8)
    sc = _abi.make_scoring(2, 4, 4, 2)
    lib = aligner._lib
    ez = np.zeros(1, dtype=_abi.RESULT_DTYPE)
    assert lib.fsv_ksw_extz2(aligner._h, 8, q.ctypes.data, 8, q.ctypes.data, 5, None, 4, 2, -1, -1, 0, 0, ez.ctypes.data, None, 0) == _abi.ERR_INVALID
    # a CIGAR that does not fit the caller's buffer: error code, result still valid (n_cigar set)
    rc, ez1, _ = _call_single(aligner, False, q, q, sc, -1, -1, 0, 0)
    assert rc == 0 and int(ez1["n_cigar"]) == 1
    mat = _abi.scoring_mat(sc)
    rc = lib.fsv_ksw_extz2(aligner._h, 8, q.ctypes.data, 8, q.ctypes.data, 5, mat.ctypes.data_as(C.POINTER(C.c_int8)), 4, 2, -1, -1, 0, 0,
                           ez.ctypes.data, None, 0)
    assert rc == _abi.ERR_CIGAR_CAP and int(ez[0]["n_cigar"]) == 1 and int(ez[0]["score"]) == 16
    # ksw2's silent return (qlen <= 0, ksw2_extz2_sse.c:57): success with a reset result
    rc, ez0, _ = _call_single(aligner, False, q[:0], q, sc, -1, -1, 0, 0)
    assert rc == 0 and int(ez0["score"]) == _abi.NEG_INF and int(ez0["max_q"]) == -1


def _ascii(codes):
    return "".join("ACGTN"[int(x)] for x in codes)


@pytest.mark.gpu
def test_c_probe_afine_gap_alignment(tmp_path, oracle):
    """gcc-built C caller with afine_gap_alignment's call sequence (Correct.cpp:7667-7697), linked against the shared object."""
    if not oracle.have_reference():
        pytest.skip("oracle/_ref/libksw2_ref.so not present")
    exe = str(tmp_path / "afine_gap_probe")
    libdir = os.path.join(ROOT, "focalsv_b200")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "c_probe", "afine_gap_probe.c"),
                           os.path.join(libdir, "libfocalsv_cuda.so"), "-Wl,-rpath," + libdir])
    rng = np.random.default_rng(77)
    from focalsv_b200.synth import mutate, random_seq
    cases = []
    SO, AM, AD, EX = _abi.EZ_SCORE_ONLY, _abi.EZ_APPROX_MAX, _abi.EZ_APPROX_DROP, _abi.EZ_EXTZ_ONLY
    # the modes hifiasm passes: exact SCORE_ONLY (Correct.cpp:7747,7799), SCORE_ONLY|APPROX_MAX|APPROX_DROP (:7811,7956), plus CIGAR modes
    modes = [0, SO, SO | AM | AD, EX, EX | SO]
    for k in range(40):
        L = int(rng.integers(20, 3000))
        t = random_seq(rng, L)
        q = mutate(rng, t, 0.03, 0.01, 0.01)
        if k % 7 == 0:
            q = q.copy(); q[rng.integers(0, len(q), 3)] = 4          # 'N' -> wildcard code 4, scored 0 by this matrix
        if k % 5 == 0 and L > 800:                                    # unrelated tail: the z-drop (400) fires
            q = np.concatenate([q[:L // 2], random_seq(rng, L // 2)])
        cases.append((_ascii(q), _ascii(t), int(k % 2), modes[k % len(modes)], [0, 2][(k // 2) % 2]))
    path = tmp_path / "cases.txt"
    path.write_text("".join("%s %s %d %d %d\n" % c for c in cases))
    out = subprocess.run([exe, str(path)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().split("\n")
    assert len(lines) == len(cases)
    sc = _abi.make_scoring(2, 4, 4, 2, sc_ambi=0)
    code = {c: i for i, c in enumerate("ACGTN")}
    for (qs, ts, strand, mode, eb), ln in zip(cases, lines):
        qn = np.array([code[c] for c in (qs if strand == 0 else qs[::-1])], dtype=np.uint8)
        tn = np.array([code[c] for c in (ts if strand == 0 else ts[::-1])], dtype=np.uint8)
        r, cg = oracle.ref_extz2(qn, tn, sc, 500, 400, eb, mode)
        want = "%d %d %d %d %d %d %d %d %d %d %s" % (r["score"], r["max"], r["mqe"], r["mqe_t"], r["mte"], r["mte_q"], r["max_t"], r["max_q"],
                                                    r["zdropped"], r["n_cigar"], _abi.cigar_str(cg))
        assert ln.strip() == want.strip()
