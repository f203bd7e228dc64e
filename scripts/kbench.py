"""Kernel-focused throughput probe: N equal tasks (no tail), one scoring, prints GCUPS per phase."""
import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np
from focalsv_b200 import api, _abi, synth
from focalsv_b200.presets import PRESETS, scoring_for
preset = sys.argv[1] if len(sys.argv)>1 else "asm5"
L = int(sys.argv[2]) if len(sys.argv)>2 else 20000
w = int(sys.argv[3]) if len(sys.argv)>3 else 3001
n = int(sys.argv[4]) if len(sys.argv)>4 else 600
flag = int(sys.argv[5],0) if len(sys.argv)>5 else 0
force = int(sys.argv[6]) if len(sys.argv)>6 else 0
opts = [a.split("=") for a in sys.argv[7:]]      # extra fsv_set_option pairs, e.g. force_excl=1
rng = np.random.default_rng(5)
pairs=[]
for i in range(n):
    ref = synth.random_seq(rng, L)
    q,_ = synth.plant_svs(rng, ref, 2, max_net=min(w//2-50, 1200), max_len=min(w//2-60, 1000))
    pairs.append((synth.mutate(rng,q,0.0006,0.0002,0.0002), ref))
g = synth._pack("k", preset, pairs, w, int(__import__("os").environ.get("KB_ZDROP", PRESETS[preset].zdrop)), flag=flag)
al = api.Aligner(0)
if force: al.set_option("force_exact",1)
for k, v in opts: al.set_option(k, int(v))
b = al.batch(g.scoring, g.qarena, g.tarena, g.tasks)
b.run()
for rep in range(3):
    b.run(); s = al.stats()
    res,_ = b.fetch(); cells = int(res["cells"].sum())
    print("%s L=%d w=%d n=%d flag=%#x: total %.1f ms fill %.1f ms bt %.1f ms  -> %.1f GCUPS (fill %.1f)  tb %.1f GB  zdropped %d" % (
        preset, L, w, n, flag, s["total_ms"], s["fill_ms"], s["backtrack_ms"], cells/s["total_ms"]/1e6, cells/s["fill_ms"]/1e6, s["traceback_bytes"]/1e9, int(res["zdropped"].sum())))
