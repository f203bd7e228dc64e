"""Re-run one case of the GPU fuzz (tests/test_gpu_parity.py::_fuzz) and print both results."""
import sys; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from oracle import oracle as O
from focalsv_b200 import api
from util import random_case, describe, same_result
seed, idx, dual = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rng = np.random.default_rng(seed)
al = api.Aligner(0)
for it in range(idx + 1):
    c = random_case(rng, max_len=400, dual=bool(dual))
fo = O.extd2 if dual else O.extz2
fg = al.extd2 if dual else al.extz2
r1, c1 = fo(c["q"], c["t"], c["sc"], w=c["w"], zdrop=c["zdrop"], end_bonus=c["end_bonus"], flag=c["flag"])
r2, c2 = fg(c["q"], c["t"], c["sc"], w=c["w"], zdrop=c["zdrop"], end_bonus=c["end_bonus"], flag=c["flag"])
print(len(c["q"]), len(c["t"]), c["w"], c["zdrop"], hex(c["flag"]), c["q"][:8], c["t"][:8])
print("oracle", describe(r1, c1), int(r1["cells"]))
print("gpu   ", describe(r2, c2), int(r2["cells"]))
