"""All five BASELINE configs (scaled by argv[1], default 0.1) through the C ABI: throughput per group and a
bit-exact spot check of a sample of tasks against the oracle."""
import sys, time; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from oracle import oracle as O
from focalsv_b200 import api, _abi, synth
from util import same_result
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
ncheck = int(sys.argv[2]) if len(sys.argv) > 2 else 6
only = sys.argv[3].split(",") if len(sys.argv) > 3 else None      # e.g. cfg3,cfg4
al = api.Aligner(0)
cfgs = {
    "cfg1": lambda: synth.config1(n_reads=max(8, int(400 * scale))),
    "cfg2": lambda: synth.config2(n_regions=max(8, int(5000 * scale)), max_region=None if scale >= 1 else 400000),
    "cfg3": lambda: synth.config3(n_regions=max(8, int(2000 * scale)), max_region=None if scale >= 1 else 200000),
    "cfg4": lambda: synth.config4(n_dup=max(4, int(213 * scale)), n_pair=max(4, int(270 * scale)), max_region=None if scale >= 1 else 1500000),
}
for name, mk in cfgs.items():
    if only and name not in only:
        continue
    t0 = time.time(); groups = mk(); tg = time.time() - t0
    for g in groups:
        b = al.batch(g.scoring, g.qarena, g.tarena, g.tasks); b.run(); b.run(); s = al.stats(); res, cig = b.fetch(); b.close()
        cells = int(res["cells"].sum()); regions = len(set(g.region_of.tolist()))
        small = np.argsort(g.tasks["qlen"].astype(np.int64) * 0 + res["cells"])[:ncheck]      # cheapest tasks for the oracle
        ores, oar = O.run_batch(g.scoring, g.qarena, g.tarena, g.tasks[small], threads=16)
        bad = 0
        for k, i in enumerate(small):
            oc = oar[int(ores[k]["cigar_off"]):int(ores[k]["cigar_off"]) + int(ores[k]["n_cigar"])]
            bad += 0 if (same_result(ores[k], oc, res[i], api.task_cigar(res[i], cig)) and int(ores[k]["cells"]) == int(res[i]["cells"])) else 1
        print("%-26s tasks %5d regions %5d cells %.3e  %.1f ms  %.1f GCUPS  %.0f regions/s  zdropped %d  general-kernel tasks %d  parity %d/%d ok" % (
            g.name, len(g.tasks), regions, cells, s["total_ms"], cells / s["total_ms"] / 1e6, regions / s["total_ms"] * 1e3,
            int(res["zdropped"].sum()), int((al.stats()["exact_path_tasks"])), len(small) - bad, len(small)),
            " segmented %d (fallbacks %d)" % (s["segmented_tasks"], s["segment_fallbacks"]), flush=True)
