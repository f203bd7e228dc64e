"""Throughput probe of fsv_edit_distance_batch on INS-allele-shaped pairs (50 bp - 12.6 kb, the chr21 truth set's
range): matrix cells (|a| x |b|) per second, against the oracle's DP on one host thread for a sample."""
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np
from focalsv_b200 import api
from oracle import oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
rng = np.random.default_rng(7)
lens = np.minimum(12600, np.maximum(50, (rng.lognormal(5.5, 1.1, n)).astype(int)))
A = [rng.integers(0, 4, int(L)).astype(np.uint8) for L in lens]
B = []
for a in A:
    b = a.copy(); k = rng.integers(0, len(b), max(1, len(b) // 20)); b[k] = rng.integers(0, 4, len(k))
    B.append(b[: max(1, len(b) - int(rng.integers(0, max(2, len(b) // 10))))])
cells = sum(len(a) * len(b) for a, b in zip(A, B))
al = api.Aligner(0)
al.edit_distances(A[:8], B[:8])
t0 = time.perf_counter(); d = al.edit_distances(A, B); t1 = time.perf_counter()
ms = al.stats()["total_ms"]
print("pairs %d  cells %.3e  kernel %.1f ms -> %.1f Gcells/s   end to end %.1f ms" % (n, cells, ms, cells / ms / 1e6, (t1 - t0) * 1e3))
idx = np.argsort(lens)[-40:]
t0 = time.perf_counter(); w = [O.edit_distance(A[i], B[i]) for i in idx]; t1 = time.perf_counter()
c2 = sum(len(A[i]) * len(B[i]) for i in idx)
print("oracle sample: %d pairs %.3e cells %.2f s -> %.2f Gcells/s (1 thread); equal: %s" % (len(idx), c2, t1 - t0, c2 / (t1 - t0) / 1e9, [int(d[i]) for i in idx] == w))
