"""Medium-size parity/determinism check of the DPX kernel (NW=6 class): TB vs score-only vs oracle."""
import sys, time; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from oracle import oracle as O
from focalsv_b200 import api, _abi, synth
from focalsv_b200.presets import PRESETS
from util import same_result, describe
preset = sys.argv[1] if len(sys.argv)>1 else "asm5"
L = int(sys.argv[2]) if len(sys.argv)>2 else 20000
w = int(sys.argv[3]) if len(sys.argv)>3 else 3001
n = int(sys.argv[4]) if len(sys.argv)>4 else 592
ncheck = int(sys.argv[5]) if len(sys.argv)>5 else 32
rng = np.random.default_rng(5)
pairs=[]
for i in range(n):
    ref = synth.random_seq(rng, L)
    q,_ = synth.plant_svs(rng, ref, 2, max_net=min(w//2-50, 1200), max_len=min(w//2-60, 1000))
    pairs.append((synth.mutate(rng,q,0.0006,0.0002,0.0002), ref))
al = api.Aligner(0)
F = ("max","zdropped","max_q","max_t","mqe","mqe_t","mte","mte_q","score","cells")
outs=[]
for flag in (0, 1, 0, 1):
    g = synth._pack("k", preset, pairs, w, PRESETS[preset].zdrop, flag=flag)
    res, cig = al.align_batch(g.scoring, g.qarena, g.tarena, g.tasks)
    outs.append((res, cig))
base = outs[0][0]
for i,(res,cig) in enumerate(outs[1:],1):
    diff = [k for k in range(n) if any(int(res[k][f])!=int(base[k][f]) for f in F)]
    print("run",i,"differs from run 0 in tasks:", diff[:10], "count", len(diff))
g = synth._pack("k", preset, pairs, w, PRESETS[preset].zdrop, flag=0)
idx = list(range(ncheck)) + [k for k in range(n) if int(base[k]["zdropped"])][:8]
sub = g.tasks[idx]
t0=time.time(); ores, oarena = O.run_batch(g.scoring, g.qarena, g.tarena, sub, threads=16); print("oracle secs", time.time()-t0)
bad=0
for j,k in enumerate(idx):
    oc = oarena[int(ores[j]["cigar_off"]):int(ores[j]["cigar_off"])+int(ores[j]["n_cigar"])]
    gc = api.task_cigar(outs[0][0][k], outs[0][1])
    if not same_result(ores[j], oc, outs[0][0][k], gc) or int(ores[j]["cells"])!=int(outs[0][0][k]["cells"]):
        bad+=1; print("BAD task",k); print(" o",describe(ores[j],oc)); print(" g",describe(outs[0][0][k],gc))
    for r_ in (1,3):
        if any(int(outs[r_][0][k][f])!=int(ores[j][f]) for f in F): print("score-only run",r_,"task",k,"differs from oracle:", {f:(int(outs[r_][0][k][f]),int(ores[j][f])) for f in F if int(outs[r_][0][k][f])!=int(ores[j][f])})
print("checked",len(idx),"bad",bad)
