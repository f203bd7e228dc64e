"""Sweep of the segmentation tunables on one workload: one shard of cfg2 (default), or cfg3 / cfg4 groups; per setting the
device time of the batch (best of 2 after one warm-up run), segmented tasks and repaired segments.
  python scripts/segsweep.py <cfg2|cfg2.r0w2|cfg3|cfg4|cfg1> "segment_auto_pct=35;segment_auto_pct=60,segment_warm_pct=300;..." """
import sys; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/scripts')
import numpy as np
import parity_full as PF
from focalsv_b200 import api, _abi
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
settings = [dict(kv.split("=") for kv in s.split(",") if kv) for s in (sys.argv[2] if len(sys.argv) > 2 else "").split(";")]
groups = PF.groups_of(cfg)
for st in settings:
    al = api.Aligner(0)
    for k, v in st.items(): al.set_option(k, int(v))
    for g in groups:
        b = al.batch(g.scoring, g.qarena, g.tarena, g.tasks)
        ms = []
        for rep in range(3):
            b.run(); ms.append(al.stats()["total_ms"])
        s = al.stats(); res, _ = b.fetch(); cells = float(res["cells"].sum())
        plan = b.plan(); nseg = int(((plan >> 16) & 0x7fff).sum())
        print("%-22s %-50s ms %s -> %.1f GCUPS | segmented %d tasks / %d segments, repaired %d" % (
            g.name, st, ["%.0f" % x for x in ms], cells / min(ms[1:]) / 1e6, s["segmented_tasks"], nseg, s["segment_fallbacks"]), flush=True)
        b.close()
    al.close()
