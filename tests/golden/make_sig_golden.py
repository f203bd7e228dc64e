"""Golden vectors for the CIGAR -> DEL/INS signature step, produced by the REFERENCE'S OWN function.

Runs in the build container only (needs /root/reference): imports
focalsv/4_sv_calling/Dippav/extract_contig_signature_CCS.py with its unavailable imports (pysam, matplotlib,
tqdm, utils) stubbed — `extract_sig_from_cigar` (lines 14-127) itself uses none of them — feeds it seeded random
alignment records and freezes inputs and outputs in tests/golden/sig_golden.json.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference/focalsv/4_sv_calling/Dippav/extract_contig_signature_CCS.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sig_golden.json")


def load_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "pysam", "tqdm", "utils"):
        m = types.ModuleType(name)
        m.tqdm = lambda x, *a, **k: x
        m.load_contigs = lambda *a, **k: None
        sys.modules.setdefault(name, m)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    spec = importlib.util.spec_from_file_location("ref_extract_sig", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class Read(object):
    def __init__(self, d):
        self.reference_name = d["reference_name"]; self.pos = d["pos"]; self.cigar = [tuple(x) for x in d["cigar"]]
        self.qname = d["qname"]; self.is_reverse = d["is_reverse"]; self.mapq = d["mapq"]


def random_record(rng, k):
    cig = []
    if rng.random() < 0.25:
        cig.append((5 if rng.random() < 0.5 else 4, int(rng.integers(1, 3000))))
    n = int(rng.integers(1, 40))
    for i in range(n):
        cig.append((0, int(rng.integers(1, 400 if rng.random() < 0.7 else 5000))))
        if i + 1 < n:
            op = 1 if rng.random() < 0.5 else 2
            big = rng.random() < 0.6
            cig.append((op, int(rng.integers(20, 1500)) if big else int(rng.integers(1, 40))))
            if rng.random() < 0.2:      # adjacent I and D, as ksw2 emits around complex events
                cig.append((3 - op, int(rng.integers(20, 800))))
    if rng.random() < 0.2:
        cig.append((4, int(rng.integers(1, 2000))))
    return {"reference_name": "chr%d" % (1 + k % 22), "pos": int(rng.integers(0, 10 ** 8)), "cigar": cig,
            "qname": "contig_%d" % k, "is_reverse": bool(rng.random() < 0.3), "mapq": int(rng.choice([0, 20, 60]))}


def main():
    ref = load_reference()
    rng = np.random.default_rng(20261018)
    cases = []
    for k in range(120):
        rec = random_record(rng, k)
        min_svlen = int(rng.choice([30, 50, 20]))
        d, i, ro, co = ref.extract_sig_from_cigar(Read(rec), min_svlen)
        cases.append({"record": rec, "min_svlen": min_svlen, "dels": d, "inss": i, "ref_end": int(ro), "contig_end": int(co)})
    with open(OUT, "w") as fh:
        json.dump({"source": REF + ":14-127", "cases": cases}, fh)
    print("wrote", OUT, len(cases), "cases")


if __name__ == "__main__":
    main()
