"""Debug: segmented run on (a) identical sequences, (b) mutated without SVs, (c) with SVs; counts boundary mismatches (needs FSV_SEG_DEBUG build)."""
import sys; sys.path.insert(0, '/root/repo')
import numpy as np
from focalsv_b200 import api, synth
from focalsv_b200.presets import PRESETS
mode = sys.argv[1]; L = int(sys.argv[2]); w = int(sys.argv[3])
rng = np.random.default_rng(5)
ref = synth.random_seq(rng, L)
if mode == "ident": q = ref.copy()
elif mode == "mut": q = synth.mutate(rng, ref, 0.0008, 0.0002, 0.0002)
else:
    q, _ = synth.plant_svs(rng, ref, L // 15000, max_net=1200, max_len=1000); q = synth.mutate(rng, q, 0.0008, 0.0002, 0.0002)
g = synth._pack("p", "asm5", [(q, ref)], w, 200)
al = api.Aligner(0); al.set_option("segment_min_diags", 1)
b = al.batch(g.scoring, g.qarena, g.tarena, g.tasks); b.run(); r, c = b.fetch()
print(mode, "done score", int(r[0]["score"]), "zdropped", int(r[0]["zdropped"]), "ms", al.stats()["total_ms"], flush=True)
