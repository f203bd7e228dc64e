"""Segmented long tasks vs the unsegmented run of the same tasks: every ksw_extz_t field and CIGAR word must be equal.
usage: segcheck.py <preset> <len> <band> <n> [segment_rows] [flag]"""
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np
from focalsv_b200 import api, _abi, synth
from focalsv_b200.presets import PRESETS
preset = sys.argv[1] if len(sys.argv) > 1 else "asm5"
L = int(sys.argv[2]) if len(sys.argv) > 2 else 150000
w = int(sys.argv[3]) if len(sys.argv) > 3 else 3001
n = int(sys.argv[4]) if len(sys.argv) > 4 else 4
seg_rows = int(sys.argv[5]) if len(sys.argv) > 5 else 0
flag = int(sys.argv[6], 0) if len(sys.argv) > 6 else 0
extra = [a.split("=") for a in sys.argv[7:]]
rng = np.random.default_rng(11)
pairs = []
for i in range(n):
    Li = L + 1000 * i
    ref = synth.random_seq(rng, Li)
    q, _ = synth.plant_svs(rng, ref, max(2, Li // 15000), max_net=min(w // 2 - 50, 1200), max_len=min(w // 2 - 60, 1000))
    pairs.append((synth.mutate(rng, q, 0.0008, 0.0002, 0.0002), ref))
g = synth._pack("seg", preset, pairs, w, PRESETS[preset].zdrop, flag=flag)
al = api.Aligner(0)
for k, v in extra: al.set_option(k, int(v))
out = {}
for mode, minr in (("whole", 0), ("segmented", 1)):
    al.set_option("segment_min_diags", minr); al.set_option("segment_rows", seg_rows)
    b = al.batch(g.scoring, g.qarena, g.tarena, g.tasks); b.run(); b.run()
    ms = al.stats()["total_ms"]; res, cig = b.fetch(); b.close()
    out[mode] = (res, cig)
    print("%-9s %8.1f ms  %.1f GCUPS  zdropped %d  launches %d" % (mode, ms, int(res["cells"].sum()) / ms / 1e6, int(res["zdropped"].sum()), al.stats()["fill_launches"]), "segmented %d fallbacks %d" % (al.stats()["segmented_tasks"], al.stats()["segment_fallbacks"]), flush=True)
ra, ca = out["whole"]; rb, cb = out["segmented"]
bad = 0
for i in range(n):
    same = all(int(ra[i][f]) == int(rb[i][f]) for f in _abi.EZ_FIELDS) and int(ra[i]["cells"]) == int(rb[i]["cells"]) and \
           np.array_equal(api.task_cigar(ra[i], ca), api.task_cigar(rb[i], cb))
    if not same:
        bad += 1
        print("task", i, {f: (int(ra[i][f]), int(rb[i][f])) for f in _abi.EZ_FIELDS + ("cells",) if int(ra[i][f]) != int(rb[i][f])})
print("mismatching tasks:", bad, "of", n)
