"""Tight traceback pool with small pages (the tests' scenario), with the lazy-pool knobs on the command line."""
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np
from focalsv_b200 import api, synth
from focalsv_b200.presets import PRESETS
al = api.Aligner(0)
al.set_option("traceback_page_bytes", 65536); al.set_option("traceback_budget_bytes", 120 * 65536)
for kv in sys.argv[1:]:
    k, v = kv.split("="); al.set_option(k, int(v))
rng = np.random.default_rng(123); pairs = []
for L in list(rng.integers(200, 3000, 60)) + [6000, 7000]:
    ref = synth.random_seq(rng, int(L)); pairs.append((synth.mutate(rng, ref, 0.01, 0.005, 0.005), ref))
for preset, w in (("asm5", 301), ("hifiasm", 100)):
    g = synth._pack("paged." + preset, preset, pairs, w, PRESETS[preset].zdrop)
    t0 = time.time(); res, cig = al.align_batch(g.scoring, g.qarena, g.tarena, g.tasks)
    print(preset, "ok %.3f s" % (time.time() - t0), "score sum", int(res["score"].astype(np.int64).sum()), flush=True)
