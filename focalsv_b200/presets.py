"""Alignment parameters at the reference's call sites.

The reference passes only minimap2 preset strings plus `-r2k`
(DipPAV_variant_call.py:103, call_DUP_from_contigs.py:114,121, align_ins2ref.py:67);
the numbers below are minimap2 2.24's values for those presets (SURVEY
appendix B) and hifiasm's in-tree constants (Correct.h:1194-1199).
"""
from collections import namedtuple

from ._abi import make_scoring

Preset = namedtuple("Preset", "name a b q e q2 e2 zdrop zdrop_inv bw bw_long sc_ambi end_bonus")

PRESETS = {
    "asm5": Preset("asm5", 1, 19, 39, 3, 81, 1, 200, 200, 2000, 100000, 1, -1),
    "asm10": Preset("asm10", 1, 9, 16, 2, 41, 1, 200, 200, 2000, 100000, 1, -1),
    "map-hifi": Preset("map-hifi", 1, 4, 6, 2, 26, 1, 400, 200, 2000, 20000, 1, -1),
    "map-pb": Preset("map-pb", 2, 4, 4, 2, 24, 1, 400, 200, 2000, 20000, 1, -1),
    "map-ont": Preset("map-ont", 2, 4, 4, 2, 24, 1, 400, 200, 2000, 20000, 1, -1),
    # hifiasm's in-tree single-affine call (Correct.h:1194-1199; Correct.cpp:7670 uses 0 for N)
    "hifiasm": Preset("hifiasm", 2, 4, 4, 2, -1, -1, 400, 400, 500, 500, 0, 0),
}


def ksw_band(bw):
    """minimap2 runs ksw2 with band bw*1.5+1: -r2k -> 3001, default 500 -> 751."""
    return int(bw * 1.5 + 1.0)


def scoring_for(name):
    p = PRESETS[name]
    return make_scoring(p.a, p.b, p.q, p.e, p.q2, p.e2, sc_ambi=p.sc_ambi)
