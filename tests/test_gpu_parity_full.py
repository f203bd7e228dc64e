"""Parity on the tasks the benchmark spends its time on ("identical ... on all tasks of the config").

The oracle ran over EVERY task of the timed configs in the build container (scripts/parity_full.py oracle <cfg>; the
digests are committed under tests/golden/parity_*.npz: 11 ksw_extz_t fields + cells + a hash of the CIGAR words per task);
here the same tasks go through the C ABI on the GPU, whole and segmented, and every digest must match.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import parity_full as PF  # noqa: E402

from focalsv_b200 import api  # noqa: E402


def _have(cfg):
    return os.path.exists(PF.path_of(cfg, 0))


def test_digests_cover_the_timed_configs():
    """host only: the digest files exist, hold one row per task and the task shapes of the workload they claim."""
    for cfg, n_groups in (("cfg1", 2), ("cfg2", 1), ("cfg3", 2), ("long1m", 1)):
        for gi in range(n_groups):
            z = np.load(PF.path_of(cfg, gi))
            assert z["digest"].shape[1] == len(PF.FIELDS) + 4 and z["digest"].shape[0] == z["cigar_hash"].shape[0] > 0
    z = np.load(PF.path_of("cfg2", 0))
    assert z["digest"].shape[0] == 10000                                 # BASELINE configs[1]: 5 000 regions x 2 contigs
    n_diag = z["digest"][:, len(PF.FIELDS)] + z["digest"][:, len(PF.FIELDS) + 1] - 1
    assert n_diag.max() > 2000000                                         # the 1.1 Mb regions are in there


@pytest.mark.gpu
@pytest.mark.parametrize("opts", [{}, {"segment_min_diags": 0}], ids=["auto-segmented", "whole"])
def test_cfg1_200kb_asm5_pairs_and_reads(opts):
    """cfg1: the two 200 kb x 200 kb asm5 contig tasks (1.1e9 cells each) and the 400 map-hifi reads, vs the oracle."""
    al = api.Aligner(0)
    for k, v in opts.items():
        al.set_option(k, v)
    try:
        assert PF.check_gpu(al, "cfg1") == 0
    finally:
        al.close()


@pytest.mark.gpu
@pytest.mark.parametrize("opts", [{}, {"segment_min_diags": 0}], ids=["auto-segmented", "whole"])
def test_1mb_asm10_pairs(opts):
    """Two 1 Mb x 1 Mb asm10 pairs (global and EXTZ_ONLY; 2 M antidiagonals, 6e9 cells each) vs the oracle, segmented and whole."""
    al = api.Aligner(0)
    for k, v in opts.items():
        al.set_option(k, v)
    try:
        assert PF.check_gpu(al, "long1m") == 0
        st = al.stats()
        assert (st["segmented_tasks"] > 0) == (not opts)
    finally:
        al.close()


@pytest.mark.gpu
def test_cfg2_every_task_of_the_benchmark_workload():
    """BASELINE configs[1] at full size, exactly bench.py's N=1 workload: all 10 000 tasks (1.58e12 cells) vs the oracle digests."""
    al = api.Aligner(0)
    try:
        assert PF.check_gpu(al, "cfg2") == 0
    finally:
        al.close()


@pytest.mark.gpu
def test_cfg3_every_task():
    """BASELINE configs[2] at full size: 4 000 single-affine band-500 contig tasks (pinned kernel) + 2 000 map-ont reads."""
    al = api.Aligner(0)
    try:
        assert PF.check_gpu(al, "cfg3") == 0
    finally:
        al.close()
