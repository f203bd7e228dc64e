"""Why is one rank of a multi-GPU job slower?  Runs every rank's shard of a W-GPU cfg2 job on ONE GPU, one after the
other (the shards are what bench.build_shard gives rank r of W), and prints per shard: device time, the plan (segmented /
exclusive tasks), and the tail of the schedule (when the last tasks ended, which they were).
  FSV_TRACE=1 python scripts/shardprobe.py <W> [ranks, e.g. 0,1] [opt=val ...]"""
import sys; sys.path.insert(0, '/root/repo')
import numpy as np, bench
from focalsv_b200 import api, _abi
W = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ranks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 and "=" not in sys.argv[2] else list(range(W))
opts = [a.split("=") for a in sys.argv[2:] if "=" in a]
bench.WORKLOAD = "cfg2"
al = api.Aligner(0)
for k, v in opts: al.set_option(k, int(v))
for r in ranks:
    g, nreg = bench.build_shard(r, W, 5000)
    nd = g.tasks["qlen"].astype(np.int64) + g.tasks["tlen"] - 1
    b = al.batch(g.scoring, g.qarena, g.tarena, g.tasks)
    plan = b.plan()
    ms = []
    for rep in range(3):
        b.run(); ms.append(al.stats()["total_ms"])
    st = al.stats(); res, _ = b.fetch()
    tl = b.timeline().astype(np.float64); t0 = tl[:, 0].min(); s = (tl[:, 0] - t0) / 1e9; e = (tl[:, 1] - t0) / 1e9
    cells = float(res["cells"].sum())
    seg = (plan & _abi.PLAN_SEGMENTED) != 0; ex = (plan & _abi.PLAN_EXCLUSIVE) != 0
    print("shard %d/%d: %d regions %d tasks, est cells %.4e real %.4e, longest %d, >400k: %d, >2M: %d | ms %s -> %.1f GCUPS | segmented %d (fallbacks %d) exclusive %d" % (
        r, W, nreg, len(g.tasks), float((nd * np.minimum(np.minimum(g.tasks["qlen"], g.tasks["tlen"]), g.tasks["w"] + 1)).sum()), cells,
        int(nd.max()), int((nd > 400000).sum()), int((nd > 2000000).sum()), ["%.0f" % x for x in ms], cells / ms[-1] / 1e6,
        int(seg.sum()), st["segment_fallbacks"], int(ex.sum())), flush=True)
    # cells done over time (by task end), and the last finishers
    order = np.argsort(e)
    cum = np.cumsum(res["cells"][order].astype(np.float64)) / cells
    for q in (0.5, 0.9, 0.95, 0.99):
        print("   %2.0f %% of the cells done at %.3f s" % (q * 100, e[order][np.searchsorted(cum, q)]), end=";")
    print("  last end %.3f s" % e.max())
    for i in order[::-1][:8]:
        print("   late: task %5d antidiagonals %8d start %.3f end %.3f (%.2f us/antidiagonal) %s%s" % (
            i, nd[i], s[i], e[i], (e[i] - s[i]) / nd[i] * 1e6, "SEG x%d " % (plan[i] >> 16) if seg[i] else "", "EXCL" if ex[i] else ""))
    ts = np.linspace(0, e.max(), 16)
    print("   running tasks over time:", [int(((s <= t) & (e > t)).sum()) for t in ts], flush=True)
    b.close()
