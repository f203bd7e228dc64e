import sys, time; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from oracle import oracle as O
from focalsv_b200 import api, _abi, synth
from util import *
al = api.Aligner(0)
N = int(sys.argv[1]) if len(sys.argv)>1 else 200
maxlen = int(sys.argv[2]) if len(sys.argv)>2 else 400
def dpx_case(rng, dual, max_len):
    tl=int(rng.integers(1,max_len)); t=rng.integers(0,4,tl).astype(np.uint8)
    if rng.random()<0.2: q=rng.integers(0,4,int(rng.integers(1,max_len))).astype(np.uint8)
    else:
        q=synth.mutate(rng,t,rng.random()*0.1,rng.random()*0.05,rng.random()*0.05)
        if rng.random()<0.4 and len(q)>20:
            k=int(rng.integers(0,len(q)-10)); L=int(rng.integers(1,min(300,len(q)//2+2)))
            q=np.concatenate([q[:k],q[k+L:]]) if rng.random()<0.5 else np.concatenate([q[:k],rng.integers(0,4,L).astype(np.uint8),q[k:]])
        if len(q)==0: q=np.array([0],dtype=np.uint8)
    a=int(rng.integers(1,4)); b=int(rng.integers(1,8)); gq=int(rng.integers(1,10)); ge=int(rng.integers(1,4))
    if dual:
        gq2=int(gq+rng.integers(0,30)); ge2=int(max(1,ge-rng.integers(0,2)))
        if rng.random()<0.2: gq,gq2,ge,ge2=gq2,gq,ge2,ge
        sc=_abi.make_scoring(a,b,gq,ge,gq2,ge2,sc_ambi=int(rng.integers(0,3)))
    else: sc=_abi.make_scoring(a,b,gq,ge,sc_ambi=int(rng.integers(0,3)))
    if rng.random()<0.3: sc = synth.scoring_for(str(rng.choice(["asm5","asm10","map-hifi","map-ont"]) if dual else "hifiasm"))
    w=int(rng.choice([-1,1,3,5,10,17,33,50,100,300,600,1200,2000]))
    zd=int(rng.choice([-1,10,50,100,400])); flag=0
    for f,pb in ((_abi.EZ_SCORE_ONLY,0.2),(_abi.EZ_EXTZ_ONLY,0.4),(_abi.EZ_REV_CIGAR,0.3)):
        if rng.random()<pb: flag|=f
    return dict(q=q,t=t,sc=sc,w=w,zdrop=zd,end_bonus=int(rng.choice([0,0,5,10,-1])),flag=flag)
for dual in (False, True):
    rng=np.random.default_rng(101+dual); nb=0; t0=time.time(); s0=al.stats()
    for it in range(N):
        c=dpx_case(rng,dual,maxlen)
        fo = O.extd2 if dual else O.extz2; fg = al.extd2 if dual else al.extz2
        r1,c1=fo(c["q"],c["t"],c["sc"],w=c["w"],zdrop=c["zdrop"],end_bonus=c["end_bonus"],flag=c["flag"])
        r2,c2=fg(c["q"],c["t"],c["sc"],w=c["w"],zdrop=c["zdrop"],end_bonus=c["end_bonus"],flag=c["flag"])
        if not same_result(r1,c1,r2,c2) or int(r1["cells"])!=int(r2["cells"]):
            nb+=1
            if nb<=4:
                sc=c["sc"]; print("BAD",it,"ql",len(c["q"]),"tl",len(c["t"]),"w",c["w"],"zd",c["zdrop"],hex(c["flag"]),"eb",c["end_bonus"],"sc",sc.mat[0],sc.mat[1],sc.q,sc.e,sc.q2,sc.e2); print("  o",describe(r1,c1), r1["cells"]); print("  g",describe(r2,c2), r2["cells"])
    s1=al.stats()
    print("dual",dual,"bad",nb,"of",N,"secs %.1f"%(time.time()-t0),"exact_path",s1["exact_path_tasks"]-s0["exact_path_tasks"])
