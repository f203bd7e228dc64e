"""Schedule analysis of one cfg2 batch: when did each task run?"""
import sys; sys.path.insert(0,'/root/repo')
import numpy as np, bench
from focalsv_b200 import api
n = int(sys.argv[1]) if len(sys.argv)>1 else 5000
g,_ = bench.build_shard(0,1,n)
al = api.Aligner(0)
for kv in sys.argv[2:]:
    k,v = kv.split("="); al.set_option(k,int(v))
b = al.batch(g.scoring,g.qarena,g.tarena,g.tasks); b.run(); b.run()
st = al.stats(); tl = b.timeline().astype(np.float64); t0 = tl[:,0].min()
s = (tl[:,0]-t0)/1e9; e = (tl[:,1]-t0)/1e9; L = g.tasks["tlen"]
print("total_ms", st["total_ms"], "last end %.3f" % e.max())
order = np.argsort(-L)
print(" rank   tlen    start    end    dur  us/diag")
for k in list(range(0,24))+[30,40,60,100,200,400]:
    i = order[k]; nd = int(g.tasks["qlen"][i])+int(L[i])-1
    print("%5d %7d %7.3f %7.3f %6.3f %6.2f" % (k, L[i], s[i], e[i], e[i]-s[i], (e[i]-s[i])/nd*1e6))
# concurrency over time
ts = np.linspace(0, e.max(), 24)
print("running tasks over time:", [int(((s<=t)&(e>t)).sum()) for t in ts])
