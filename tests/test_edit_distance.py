"""Next row f3: batched global edit distance (what the reference asks edlib for when it de-duplicates INS alleles,
focalsv/4_sv_calling/Dippav/remove_redundancy.py:57-63: edlib.align(seq1, seq2)["editDistance"], default mode NW).
The value is unique, so the oracle is the textbook DP and parity is exact."""
import numpy as np
import pytest

from focalsv_b200 import _abi
from focalsv_b200.api import FsvError


def _py_levenshtein(a, b):
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def test_oracle_known_answers(oracle):
    for a, b, d in (("kitten", "sitting", 3), ("flaw", "lawn", 2), ("intention", "execution", 5), ("", "ACGT", 4),
                    ("ACGT", "", 4), ("ACGTACGT", "ACGTACGT", 0), ("AAAA", "TTTTTT", 6), ("GATTACA", "GCATGCU", 4)):
        assert oracle.edit_distance(a, b) == d
    rng = np.random.default_rng(4)
    for _ in range(40):
        a = "".join(rng.choice(list("ACGTN"), int(rng.integers(0, 60))))
        b = "".join(rng.choice(list("ACGTN"), int(rng.integers(0, 60))))
        assert oracle.edit_distance(a, b) == _py_levenshtein(a, b) == oracle.edit_distance(b, a)


def _mutated(rng, s, rate):
    out = []
    for c in s:
        r = rng.random()
        if r < rate / 3:
            continue
        if r < 2 * rate / 3:
            out.append(int(rng.integers(0, 4)))
        elif r < rate:
            out.append(int(c)); out.append(int(rng.integers(0, 4)))
            continue
        else:
            out.append(int(c))
    return np.array(out, dtype=np.uint8)


@pytest.mark.gpu
def test_gpu_edit_distance_equals_the_oracle(oracle, aligner):
    """Lengths around the 64-row block and 2048-row strip boundaries, multi-strip patterns, near-identical and
    unrelated pairs, ASCII and code alphabets."""
    rng = np.random.default_rng(2026)
    A, B = [], []
    for la in (1, 2, 63, 64, 65, 127, 128, 129, 500, 2047, 2048, 2049, 4096, 5000, 6500):
        a = rng.integers(0, 4, la).astype(np.uint8)
        for kind in range(3):
            if kind == 0:
                b = _mutated(rng, a, 0.08)
            elif kind == 1:
                b = rng.integers(0, 4, int(rng.integers(1, la + 200))).astype(np.uint8)
            else:
                b = np.concatenate([a[: la // 2], rng.integers(0, 4, int(rng.integers(0, 300))).astype(np.uint8), a[la // 2:]])
            if len(b) == 0:
                b = np.array([1], dtype=np.uint8)
            A.append(a); B.append(b)
    A += ["ACGTNNACGT", "", "ACGT", b"GATTACA", "kitten"]
    B += ["ACGTACGT", "ACGT", "", b"GCATGCT", "sitten"]
    # one ASCII batch and one code batch (at most 8 distinct byte values per call)
    got_codes = aligner.edit_distances(A[:-5], B[:-5])
    got_ascii = aligner.edit_distances(A[-5:-1], B[-5:-1])
    want = [oracle.edit_distance(a, b) for a, b in zip(A, B)]
    assert list(got_codes) == want[:-5]
    assert list(got_ascii) == want[-5:-1]
    assert list(aligner.edit_distances(A[-1:], B[-1:])) == want[-1:] == [1]
    assert list(aligner.edit_distances(B[:-5], A[:-5])) == want[:-5]          # symmetric


@pytest.mark.gpu
def test_gpu_edit_distance_rejects_large_alphabets(aligner):
    with pytest.raises(FsvError) as ei:
        aligner.edit_distances(["ABCDEFGHIJ"], ["JIHGFEDCBA"])
    assert ei.value.code == _abi.ERR_INVALID
    assert len(aligner.edit_distances([], [])) == 0
