import sys, time; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from oracle import oracle as O
from focalsv_b200 import api, _abi, synth
from util import *
al = api.Aligner(0)
al.set_option("force_exact", 1)
for dual in (False, True):
    rng=np.random.default_rng(11); nb=0
    t0=time.time()
    for it in range(int(sys.argv[1]) if len(sys.argv)>1 else 200):
        c=random_case(rng, dual=dual)
        fo = O.extd2 if dual else O.extz2; fg = al.extd2 if dual else al.extz2
        r1,c1=fo(c["q"],c["t"],c["sc"],w=c["w"],zdrop=c["zdrop"],end_bonus=c["end_bonus"],flag=c["flag"])
        r2,c2=fg(c["q"],c["t"],c["sc"],w=c["w"],zdrop=c["zdrop"],end_bonus=c["end_bonus"],flag=c["flag"])
        if not same_result(r1,c1,r2,c2) or int(r1["cells"])!=int(r2["cells"]):
            nb+=1
            if nb<=5: print("BAD",it,len(c["q"]),len(c["t"]),c["w"],c["zdrop"],hex(c["flag"]),c["end_bonus"]); print("  o",describe(r1,c1), r1["cells"]); print("  g",describe(r2,c2), r2["cells"])
    print("dual",dual,"bad",nb,"secs",time.time()-t0)
print(al.stats())
