"""Parity gate over EVERY task of a timed config ("identical ... on all tasks of the config").

  python scripts/parity_full.py oracle <cfg> [threads]   # CPU only (no GPU needed): runs the oracle over every task and
                                                         # writes the per-task digests to tests/golden/parity_<cfg>.npz
  python scripts/parity_full.py gpu <cfg> [opt=val ...]  # GPU box: runs the same tasks through the C ABI and compares
                                                         # every ksw_extz_t field, `cells` and the CIGAR words (by hash)
                                                         # with the committed digests; exit code 1 on any difference

<cfg>: cfg2 (bench.py's N=1 workload, exactly bench.build_shard(0, 1, 5000)), cfg2.r<R>w<W> (rank R's shard of a W-GPU
job), cfg1, cfg3, cfg4 (synth.config1/3/4 at full size).  The oracle side runs where there is no GPU (the build
container, 8 host threads: cfg2 takes about 12 minutes) so that GPU-box minutes go to the GPU; the digests are small
(64 bytes per task) and are committed next to the script that made them.  A digest = the 11 ksw_extz_t fields + cells +
blake2b-64 of the CIGAR words + (qlen, tlen, w, flag) of the task to catch a workload that drifted.
"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from focalsv_b200 import _abi, synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
FIELDS = _abi.EZ_FIELDS + ("cells",)


def groups_of(cfg):
    if cfg.startswith("cfg2"):
        import bench
        rank, world = 0, 1
        if "." in cfg:
            tag = cfg.split(".")[1]
            rank, world = int(tag[1:tag.index("w")]), int(tag[tag.index("w") + 1:])
        bench.WORKLOAD = "cfg2"
        g, _ = bench.build_shard(rank, world, 5000)
        return [g]
    if cfg == "cfg1":
        return synth.config1()
    if cfg == "cfg3":
        return synth.config3()
    if cfg == "cfg4":
        return synth.config4()
    if cfg == "long1m":
        # two 1 Mb x 1 Mb asm10 pairs (a global task and an extension), the shape of cfg4's DUP windows: 2 M antidiagonals each,
        # always segmented in auto mode
        from focalsv_b200.presets import PRESETS, ksw_band
        rng = np.random.default_rng(4242)
        pairs = []
        for k in range(2):
            ref = synth.random_seq(rng, 1000000)
            # task 0: a few small SVs, the alignment runs to the end; task 1: many large ones, ksw2's z-drop ends it inside a later segment
            q, _ = synth.plant_svs(rng, ref, 60, max_net=1200, max_len=1000) if k else synth.plant_svs(rng, ref, 12, max_net=300, max_len=150)
            pairs.append((synth.mutate(rng, q, 0.003, 0.001, 0.001), ref))
        p = PRESETS["asm10"]
        return [synth._pack("long1m.asm10", "asm10", pairs, ksw_band(p.bw), p.zdrop, flags=np.array([0, _abi.EZ_EXTZ_ONLY], dtype=np.int32))]
    raise SystemExit("unknown config %r" % cfg)


def digest(tasks, res, arena):
    n = len(tasks)
    d = np.zeros((n, len(FIELDS) + 4), dtype=np.int64)
    for k, f in enumerate(FIELDS):
        d[:, k] = res[f]
    d[:, len(FIELDS) + 0] = tasks["qlen"]; d[:, len(FIELDS) + 1] = tasks["tlen"]
    d[:, len(FIELDS) + 2] = tasks["w"]; d[:, len(FIELDS) + 3] = tasks["flag"]
    h = np.zeros(n, dtype=np.uint64)
    off = res["cigar_off"].astype(np.int64); nc = res["n_cigar"].astype(np.int64)
    buf = np.ascontiguousarray(arena, dtype=np.uint32)
    for i in range(n):
        h[i] = int.from_bytes(hashlib.blake2b(buf[off[i]:off[i] + nc[i]].tobytes(), digest_size=8).digest(), "little")
    return d, h


def path_of(cfg, gi):
    return os.path.join(GOLDEN, "parity_%s_g%d.npz" % (cfg, gi))


def main():
    mode, cfg = sys.argv[1], sys.argv[2]
    t0 = time.time()
    groups = groups_of(cfg)
    print("# %s: %d group(s), synthesised in %.1f s" % (cfg, len(groups), time.time() - t0), flush=True)
    if mode == "oracle":
        from oracle import oracle as O
        threads = int(sys.argv[3]) if len(sys.argv) > 3 else (os.cpu_count() or 1)
        for gi, g in enumerate(groups):
            t0 = time.time()
            res, arena = O.run_batch(g.scoring, g.qarena, g.tarena, g.tasks, threads=threads)
            dt = time.time() - t0
            d, h = digest(g.tasks, res, arena)
            np.savez_compressed(path_of(cfg, gi), name=g.name, fields=np.array(FIELDS), digest=d, cigar_hash=h)
            print("%-28s %6d tasks  %.3e cells  oracle %.1f s on %d threads (%.2f GCUPS)  zdropped %d -> %s" % (
                g.name, len(g.tasks), float(res["cells"].sum()), dt, threads, res["cells"].sum() / dt / 1e9,
                int(res["zdropped"].sum()), os.path.relpath(path_of(cfg, gi), ROOT)), flush=True)
        return 0
    if mode != "gpu":
        raise SystemExit(__doc__)
    from focalsv_b200 import api
    al = api.Aligner(0)
    for kv in sys.argv[3:]:
        k, v = kv.split("=")
        al.set_option(k, int(v))
    bad_total = check_gpu(al, cfg, groups)
    al.close()
    print("# %s: %s" % (cfg, "ALL TASKS BIT-EXACT vs the oracle digests" if bad_total == 0 else "%d MISMATCHES" % bad_total))
    return 1 if bad_total else 0


def check_gpu(al, cfg, groups=None, log=print):
    """Every task of `cfg` through the C ABI on aligner `al`, compared with the committed oracle digests.  Returns the
    number of mismatching tasks (a workload that drifted from the digest file counts as one)."""
    groups = groups_of(cfg) if groups is None else groups
    bad_total = 0
    for gi, g in enumerate(groups):
        z = np.load(path_of(cfg, gi))
        b = al.batch(g.scoring, g.qarena, g.tarena, g.tasks)
        plan = b.plan()
        b.run()
        s = al.stats()
        res, arena = b.fetch()
        b.close()
        d, h = digest(g.tasks, res, arena)
        od, oh = z["digest"], z["cigar_hash"]
        if od.shape != d.shape:
            log("%-28s WORKLOAD DRIFT: %s tasks here, %s in the digest file" % (g.name, d.shape, od.shape)); bad_total += 1; continue
        nf = len(FIELDS)
        if not np.array_equal(od[:, nf:], d[:, nf:]):
            log("%-28s WORKLOAD DRIFT: task shapes differ from the digest file" % g.name); bad_total += 1; continue
        bad = np.flatnonzero((od[:, :nf] != d[:, :nf]).any(axis=1) | (oh != h))
        n_diag = g.tasks["qlen"].astype(np.int64) + g.tasks["tlen"] - 1
        log("%-28s %6d tasks (longest %d antidiagonals, %d tasks >= 400k)  %.3e cells  %.1f ms  %.1f GCUPS  segmented %d fallbacks %d  exclusive %d  general-kernel %d  zdropped %d  MISMATCHES %d" % (
            g.name, len(g.tasks), int(n_diag.max()), int((n_diag >= 400000).sum()), float(res["cells"].sum()), s["total_ms"],
            res["cells"].sum() / s["total_ms"] / 1e6, s["segmented_tasks"], s["segment_fallbacks"], int((plan & _abi.PLAN_EXCLUSIVE != 0).sum()),
            int((plan & _abi.PLAN_GENERAL != 0).sum()), int(res["zdropped"].sum()), len(bad)))
        for i in bad[:10]:
            diff = [f for k, f in enumerate(FIELDS) if od[i, k] != d[i, k]]
            log("   task %d (qlen %d tlen %d): fields %s cigar %s" % (i, g.tasks["qlen"][i], g.tasks["tlen"][i], diff, "differs" if oh[i] != h[i] else "same"))
        bad_total += len(bad)
    return bad_total


if __name__ == "__main__":
    sys.exit(main())
