#!/usr/bin/env python
"""bench.py — alignment GCUPS / target regions per second of the FocalSV alignment-DP hot path.

  python bench.py --gpus N --steps K --warmup W            # our CUDA arm (libfocalsv_cuda)
  python bench.py --impl reference --gpus N --steps K ...   # the reference arm: CPU, host cores

A "step" is one pass of the hot path (DP fill + CIGAR reconstruction) over one batch of
synthetic input: BASELINE.json configs[1] = FocalSV auto mode, 5 000 SV-rich regions x 2
haplotype contigs vs hg38-shaped windows, asm5 scoring, band 3001 (-r2k), z-drop 200
(focalsv_b200/synth.py config2, seed 1002).  Multi-GPU = weak scaling: every rank owns a
length-balanced (LPT) shard of N x 5 000 regions; regions are independent, so there is no
collective on the data path (torch.distributed only provides the barrier and max-over-ranks).

One JSON line on rank 0; see README / DESIGN.md §6 for the fields.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from focalsv_b200 import _abi, synth  # noqa: E402
from focalsv_b200.presets import PRESETS, ksw_band  # noqa: E402

# canonical integer lane-ops per in-band cell: exact max + traceback
OPS_PER_CELL = {"extz2": 35, "extd2": 54}
CONFIG_SEED = 1002


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--regions", type=int, default=int(os.environ.get("FSV_BENCH_REGIONS", 5000)),
                    help="regions per GPU (configs[1] = 5000)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work per reference step / baseline sample")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3"],
                    help="cfg2 = BASELINE configs[1] (the metric's config, default); cfg3 = BASELINE configs[2]'s single-affine "
                         "contig tasks (a=2,b=4,q=4,e=2,w=500,z=400): the one workload whose CPU arm is the reference's OWN "
                         "ksw2_extz2_sse.c compiled in place (cpu_baseline.kind = reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--parity-tasks", type=int, default=40, help="live oracle sample: tasks over the size quantiles (the longest included)")
    ap.add_argument("--parity-segmented", type=int, default=6, help="live oracle sample: extra segmented tasks (the longest of them always)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--force-exact", action="store_true", help="use only the general int8-exact kernel")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


WORKLOAD = "cfg2"      # set from --workload in main()


def build_shard(rank, world, regions_per_gpu):
    """Region lengths of the whole job (world x regions_per_gpu), LPT-binned by estimated cells;
    this rank synthesises only its own bin."""
    if WORKLOAD == "cfg3":      # every rank its own 1/world-independent share (weak scaling), single-affine contig tasks
        g = synth.config3(n_regions=regions_per_gpu, seed=1003 + 7919 * rank, max_region=400000)[0]
        return g, regions_per_gpu
    rng = np.random.default_rng(CONFIG_SEED)
    lens = synth.sample_quantiles(rng, synth.REGION_LEN_Q, regions_per_gpu * world)
    if world > 1:
        w = ksw_band(PRESETS["asm5"].bw)
        est = (2 * lens - 1) * np.minimum(lens, w + 1)
        order = np.argsort(-est, kind="stable")
        load = np.zeros(world, dtype=np.int64)
        mine = []
        for i in order:                       # longest-processing-time-first
            b = int(np.argmin(load))
            load[b] += est[i]
            if b == rank:
                mine.append(i)
        lens = lens[np.array(sorted(mine), dtype=np.int64)]
    groups = synth.config2(seed=CONFIG_SEED + 7919 * rank, lens=lens)
    return groups[0], len(lens)


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.proc, self.lines = gpu, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_sample(group, seconds, cores, gcups_guess):
    """Bounded sample of the same workload: whole tasks, largest-first stride, ~`seconds` of CPU work."""
    t = group.tasks
    est = (t["qlen"].astype(np.int64) + t["tlen"] - 1) * np.minimum(np.minimum(t["qlen"], t["tlen"]), t["w"] + 1)
    budget = seconds * cores * gcups_guess * 1e9
    rng = np.random.default_rng(99)
    pick, tot = [], 0
    for i in rng.permutation(len(t)):
        if est[i] > budget * 0.5 and pick:
            continue
        pick.append(int(i)); tot += int(est[i])
        if tot >= budget or len(pick) >= 4 * cores:
            if tot >= budget:
                break
    return np.array(sorted(pick), dtype=np.int64)


def run_cpu(group, idx, threads):
    """The oracle timed as the CPU baseline (the one place bench.py executes oracle/)."""
    from oracle import oracle as O
    sub = group.tasks[idx]
    use_ref = O.have_reference() and group.scoring.q2 < 0
    t0 = time.perf_counter()
    res, _ = O.run_batch(group.scoring, group.qarena, group.tarena, sub, threads=threads, use_reference=use_ref)
    dt = time.perf_counter() - t0
    return int(res["cells"].sum()), dt, ("reference" if use_ref else "port")


def reference_arm(args, rank, world):
    if rank != 0:
        return
    group, n_regions = build_shard(0, 1, args.regions)
    cores = os.cpu_count() or 1
    idx = cpu_sample(group, args.cpu_seconds, cores, 0.45)
    regions = len(set(group.region_of[idx].tolist()))
    for _ in range(max(args.warmup, 0)):
        run_cpu(group, idx[: max(1, len(idx) // 8)], cores)
    cells, secs, kind = 0, 0.0, "port"
    for _ in range(args.steps):
        c, dt, kind = run_cpu(group, idx, cores)
        cells += c; secs += dt
    gcups = cells / secs / 1e9
    idx1 = cpu_sample(group, max(args.cpu_seconds / 4, 1.0), 1, 0.45)
    c1, dt1, _ = run_cpu(group, idx1, 1)
    sample = "%d of %d tasks of %s (seed %d), %.3g cells/step, handed out largest first" % (len(idx), len(group.tasks), WORKLOAD, CONFIG_SEED, cells / max(args.steps, 1))
    line = {"impl": "reference", "metric": "alignment_gcups", "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / max(args.steps, 1) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "regions_per_s": regions * args.steps / secs,
            "config": workload_config(args, world, n_regions),
            "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": cores, "kind": kind, "sample": sample,
                             "t1_value": c1 / dt1 / 1e9, "t1_sample": "%d tasks (%.3g cells) on one thread" % (len(idx1), c1)},
            "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": ("the reference's own software/hifiasm-0.16.1/ksw2_extz2_sse.c, compiled in place into oracle/_ref "
                     "(gcc -O3 -msse4.1), one task per thread, all host threads") if kind == "reference" else
                    ("port, source-unpinned: dual-affine ksw_extd2_sse is not in /root/reference (minimap2 2.24 dependency); the CPU arm is the "
                     "oracle's plain-C restatement (gcc -O3 -msse4.1 + AVX2 clone), one task per thread, largest first, all host threads")}
    print(json.dumps(line), flush=True)


def workload_config(args, world, n_regions):
    if WORKLOAD == "cfg3":
        return {"workload": "BASELINE configs[2] (contig tasks): %d regions x 2 contigs per GPU, 1 %% error, single-affine a=2,b=4,q=4,e=2 "
                            "(hifiasm's in-tree constants, Correct.h:1194-1199), band 500, zdrop 400, global / EXTZ_ONLY + CIGAR; regions capped at 400 kb" % args.regions,
                "regions_per_gpu": args.regions, "regions_this_rank": n_regions, "tasks_per_region": 2, "seed": 1003,
                "sharding": "independent shares per rank, no collective", "world": world,
                "l2_policy": "inputs+traceback per step exceed the 126 MB L2"}
    return {"workload": "BASELINE configs[1]: FocalSV auto mode, %d SV-rich regions x 2 haplotype contigs per GPU vs "
                        "hg38-shaped windows, asm5 (a=1,b=19,q=39,e=3,q2=81,e2=1), band 3001, zdrop 200, global + CIGAR; planted SVs "
                        "1 per 15 kb with lengths from the chr21 truth set capped at 1.2 kb (band/2 - 300: a longer net offset leaves "
                        "a band-3001 global task by construction; SURVEY 8d samples up to 12.6 kb)" % args.regions,
            "regions_per_gpu": args.regions, "regions_this_rank": n_regions, "tasks_per_region": 2,
            "seed": CONFIG_SEED, "sharding": "LPT bins by estimated cells, no collective", "world": world,
            "l2_policy": "inputs+traceback per step far exceed the 126 MB L2 (traceback alone is >100 GB/step)"}


def main():
    args = parse_args()
    global WORKLOAD
    WORKLOAD = args.workload
    if WORKLOAD == "cfg3" and args.regions == 5000:
        args.regions = 2000                  # BASELINE configs[2]: 2 000 regions
    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from focalsv_b200 import api
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (ours) needs a CUDA device: libfocalsv_cuda has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    group, n_regions = build_shard(rank, world, args.regions)
    kind = "extd2" if group.scoring.q2 >= 0 else "extz2"
    al = api.Aligner(local_rank)
    if args.force_exact:
        al.set_option("force_exact", 1)

    # ---- device-resident arm: inputs in HBM before the clock starts
    batch = al.batch(group.scoring, group.qarena, group.tarena, group.tasks)
    batch_plan = batch.plan()
    for _ in range(args.warmup):
        batch.run()
    barrier()
    st0 = al.stats()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    dev_ms = fill_ms = bt_ms = 0.0
    for _ in range(args.steps):
        batch.run()                       # kernels on the library's stream; it synchronises that stream
        s = al.stats()
        dev_ms += s["total_ms"]; fill_ms += s["fill_ms"]; bt_ms += s["backtrack_ms"]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    st1 = al.stats()
    res, cig = batch.fetch()
    batch.close()
    cells_step = int(res["cells"].sum())
    dev_s = max_over_ranks(dev_ms * 1e-3)         # CUDA events on the launching stream, max over ranks
    # per-rank view (the job's time is the slowest rank's): device ms per step, cells, segmented tasks / fallbacks
    mine = [dev_ms / args.steps, float(cells_step), float(st1["segmented_tasks"]), float(st1["segment_fallbacks"]),
            float((batch_plan & _abi.PLAN_EXCLUSIVE != 0).sum()), float(len(group.tasks))]
    if world > 1:
        tt = torch.tensor(mine, dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(allr, tt)
        per_rank = [[float(x) for x in t.tolist()] for t in allr]
    else:
        per_rank = [mine]
    wall_s = max_over_ranks(wall)
    tot_cells = sum_over_ranks(float(cells_step)) * args.steps
    tot_regions = sum_over_ranks(float(n_regions)) * args.steps
    gcups = tot_cells / dev_s / 1e9
    launches = (st1["fill_launches"] + st1["backtrack_launches"] + st1["other_launches"]
                - st0["fill_launches"] - st0["backtrack_launches"] - st0["other_launches"])

    # ---- parity gate on what was just timed (oracle as checker only).  (1) When this is the default N=1 workload, EVERY
    # task is compared with the committed oracle digests (tests/golden/parity_cfg2_g0.npz, made by scripts/parity_full.py
    # oracle cfg2).  (2) Always: a live oracle run in this process on a sample stratified over the size quantiles that
    # includes the longest task, segmented tasks (the cold-start + stitch path) and exclusive-launch tasks.
    parity = None
    plan = batch_plan
    seg_mask = (plan & _abi.PLAN_SEGMENTED) != 0
    if rank == 0 and not args.no_parity:
        from oracle import oracle as O
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import parity_full as PF
        parity = {"bit_exact": True}
        dpath = PF.path_of(WORKLOAD, 0)
        if world == 1 and args.regions == (5000 if WORKLOAD == "cfg2" else -1) and os.path.exists(dpath):
            z = np.load(dpath)
            d, h = PF.digest(group.tasks, res, cig)
            nf = len(PF.FIELDS)
            same_work = z["digest"].shape == d.shape and np.array_equal(z["digest"][:, nf:], d[:, nf:])
            n_bad = int(((z["digest"][:, :nf] != d[:, :nf]).any(axis=1) | (z["cigar_hash"] != h)).sum()) if same_work else -1
            parity["all_tasks"] = {"tasks_checked": int(len(d)) if same_work else 0, "mismatches": n_bad,
                                   "against": "tests/golden/parity_%s_g0.npz (oracle digests of every task: 11 ksw_extz_t fields, cells, CIGAR hash)" % WORKLOAD}
            parity["bit_exact"] &= same_work and n_bad == 0
        n_diag = group.tasks["qlen"].astype(np.int64) + group.tasks["tlen"] - 1
        order = np.argsort(n_diag, kind="stable")
        rng = np.random.default_rng(7)
        pick = set(order[np.linspace(0, len(order) - 1, args.parity_tasks).astype(np.int64)].tolist())      # size quantiles, longest included
        segs = np.flatnonzero(seg_mask)
        if len(segs):
            pick.add(int(segs[np.argmax(n_diag[segs])]))
            pick.update(int(x) for x in rng.choice(segs, min(args.parity_segmented, len(segs)), replace=False))
        excl = np.flatnonzero((plan & _abi.PLAN_EXCLUSIVE) != 0)
        pick.update(int(x) for x in excl[:2])
        pick = np.array(sorted(pick), dtype=np.int64)
        t0 = time.perf_counter()
        ores, oarena = O.run_batch(group.scoring, group.qarena, group.tarena, group.tasks[pick], threads=os.cpu_count() or 1)
        ok = True
        for k, i in enumerate(pick):
            gc = api.task_cigar(res[i], cig)
            oc = oarena[int(ores[k]["cigar_off"]):int(ores[k]["cigar_off"]) + int(ores[k]["n_cigar"])]
            ok &= all(int(res[i][f]) == int(ores[k][f]) for f in _abi.EZ_FIELDS) and np.array_equal(gc, oc) and int(res[i]["cells"]) == int(ores[k]["cells"])
        parity["live_sample"] = {"tasks_checked": int(len(pick)), "bit_exact": bool(ok), "longest_antidiagonals": int(n_diag[pick].max()),
                                 "segmented_in_sample": int(seg_mask[pick].sum()), "exclusive_in_sample": int(((plan[pick] & _abi.PLAN_EXCLUSIVE) != 0).sum()),
                                 "oracle_seconds": time.perf_counter() - t0}
        parity["bit_exact"] = bool(parity["bit_exact"] and ok)
        parity["tasks_checked"] = int(max(len(pick), parity.get("all_tasks", {}).get("tasks_checked", 0)))

    # ---- end-to-end arm: host buffers through the public API, H2D and D2H inside the clock
    e2e = None
    if not args.no_e2e:
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
        q_p, t_p = pin(group.qarena), pin(group.tarena)
        tasks_p = pin(group.tasks.view(np.uint8)).view(_abi.TASK_DTYPE)
        out_p = torch.empty(len(group.tasks) * _abi.RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory().numpy().view(_abi.RESULT_DTYPE)
        cig_p = torch.empty(max(len(cig) + 1024, 1024), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
        al.align_batch(group.scoring, q_p, t_p, tasks_p, out=out_p, cig=cig_p)       # warm
        barrier()
        s0 = al.stats()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r2, c2 = al.align_batch(group.scoring, q_p, t_p, tasks_p, out=out_p, cig=cig_p)
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        s1 = al.stats()
        e2e = {"value": tot_cells / e2e_s / 1e9, "unit": "GCUPS",
               "h2d_bytes_per_step": (s1["h2d_bytes"] - s0["h2d_bytes"]) // args.steps,
               "d2h_bytes_per_step": (s1["d2h_bytes"] - s0["d2h_bytes"]) // args.steps,
               "regions_per_s": tot_regions / e2e_s, "ms_per_step": e2e_s / args.steps * 1e3,
               "timing": "host wall clock around fsv_align_batch (pinned host buffers; H2D + kernels + D2H), max over ranks"}

    # ---- roofline of the dominant kernel (DP fill): integer issue rate, measured on this device
    peaks = {}
    names = {0: "VIADD.16x2", 1: "VIMNMX3.S16x2", 2: "VIADDMNMX.S16x2", 3: "LOP3", 4: "IMAD", 5: "fill-mix", 6: "PRMT", 7: "VIMNMX3+IMAD"}
    for k, nm in names.items():
        peaks[nm] = al.int_peak(k) / 1e12
    peak = max(peaks.values())
    fill_s = max_over_ranks(fill_ms * 1e-3)
    achieved = cells_step * args.steps * OPS_PER_CELL[kind] / (fill_ms * 1e-3) / 1e12
    hbm_peak = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            hbm_peak = json.load(fh).get("hbm_gbs")
    except Exception:
        pass
    tb_gbs = st1["traceback_bytes"] * args.steps / (fill_ms * 1e-3) / 1e9
    # DRAM bytes of the fill launches of one step of THIS workload, from the committed ncu capture (profiles/);
    # scaled by cells when the run uses another number of regions
    traffic = None
    try:
        if WORKLOAD != "cfg2":
            raise KeyError("the committed capture is of cfg2")
        with open(os.path.join(ROOT, "profiles", "ncu_traffic_r2.json")) as fh:
            tj = json.load(fh)
        traffic = {"dram_bytes_per_step": tj["dram_bytes_per_step"] * (cells_step / tj["cells_per_step"]),
                   "algorithmic_bytes_per_step": float(cells_step), "unit": "bytes (1 B traceback per in-band cell)",
                   "source": "profiles/ncu_traffic_r2.json (ncu dram__bytes_read.sum + dram__bytes_write.sum of the fill launches of one step, round-2 build: " + ", ".join(sorted(tj["launches"])) + ")"}
    except Exception:
        pass
    # Two views of the same kernel.  (1) SURVEY 8(d)'s contract: canonical SSE lane-ops per cell (54 dual-affine, 35 single-affine, exact maximum +
    # traceback) x cells / fill time, against the measured peak IN THE SAME UNIT: the fastest dependency-free stream found (VIMNMX3.S16x2 and IMAD
    # alternating on the two integer pipes) retires 1.5 canonical ops per instruction (a 3-input max is two of the reference's max ops), so
    # peak = 1.5 x its measured instruction rate.  (Round 1 divided by the instruction rate itself; the DPX kernel fuses ops, so that fraction passes 1.)
    # (2) what the hardware sees: thread-instructions actually issued per cell (ncu, profiles/ncu_dpx_fill_r2.json) x cells / fill time against the
    # measured dual-pipe instruction rate, with ncu's issue-slot and ALU-pipe utilisation of the same capture beside it.
    canon_per_instr = 1.5
    issue = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_dpx_fill_r2.json")) as fh:
            nj = json.load(fh)
        if kind == "extd2":
            ach_i = cells_step * args.steps * nj["thread_instructions_per_cell"] / (fill_ms * 1e-3) / 1e12
            issue = {"achieved": ach_i, "peak": peak, "unit": "Tlane-instr/s", "frac": ach_i / peak,
                     "thread_instructions_per_cell": nj["thread_instructions_per_cell"],
                     "ncu": {"issue_active_pct": nj["issue_active_pct"], "alu_pipe_pct": nj["alu_pipe_pct"], "fma_pipe_pct": nj["fma_pipe_pct"],
                             "barrier_stall_per_issue": nj["barrier_stall_per_issue"], "registers_per_thread": nj["registers_per_thread"]},
                     "source": "profiles/ncu_dpx_fill_r2.json (one launch of %s on uniform band-3001 tasks under ncu --set full)" % nj["kernel"]}
    except Exception:
        pass
    roofline = {"bound": "int_alu", "kernel": "fsv_fill (DP fill incl. traceback store)",
                "achieved": achieved, "peak": peak * canon_per_instr, "unit": "T canonical lane-op/s (SURVEY 8d)", "frac": achieved / (peak * canon_per_instr),
                "ops_per_cell": OPS_PER_CELL[kind], "cells_per_launch": cells_step / max(1, (st1["fill_launches"] - st0["fill_launches"]) // max(args.steps, 1)),
                "peak_source": "1.5 canonical ops per instruction x the best dependency-free instruction rate measured in this run (fsv_measure_int_peak, Tlane-instr/s): "
                               + ", ".join("%s=%.1f" % kv for kv in sorted(peaks.items())),
                "issue": issue,
                "traffic": traffic,
                "hbm": {"achieved": tb_gbs, "peak": hbm_peak or 6650.0, "unit": "GB/s",
                        "frac": tb_gbs / (hbm_peak or 6650.0), "what": "traceback bytes written / fill time",
                        "peak_source": "MEASURED_PEAKS.json" if hbm_peak else "fallback"},
                "fill_share_of_step": fill_s / dev_s if dev_s else None,
                "backtrack_ms_per_step": bt_ms / args.steps}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        idx = cpu_sample(group, args.cpu_seconds, cores, 0.45)
        c, dt, knd = run_cpu(group, idx, cores)
        idx1 = cpu_sample(group, max(args.cpu_seconds / 4, 1.0), 1, 0.45)
        c1, dt1, _ = run_cpu(group, idx1, 1)
        cpu_baseline = {"value": c / dt / 1e9, "unit": "GCUPS", "cores": cores, "kind": knd,
                        "sample": "%d of %d tasks (%.3g cells) of the same workload, largest first, one task per thread" % (len(idx), len(group.tasks), c),
                        "seconds": dt, "t1_value": c1 / dt1 / 1e9, "t1_sample": "%d tasks (%.3g cells) on one thread" % (len(idx1), c1),
                        "note": "port = the oracle's plain-C restatement (dual-affine source unpinned: minimap2 2.24 is not in the reference tree)" if knd == "port" else "the reference's own ksw2_extz2_sse.c compiled in place"}

    if rank == 0:
        line = {"metric": "alignment_gcups", "value": gcups, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
                "regions_per_s": tot_regions / dev_s, "wall_ms_per_step": wall_s / args.steps * 1e3,
                "config": workload_config(args, world, n_regions), "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
                "parity": parity, "exact_path_tasks": int(st1["exact_path_tasks"] - st0["exact_path_tasks"]) // max(args.steps, 1),
                "segmented_tasks": int(st1["segmented_tasks"]), "segment_fallbacks": int(st1["segment_fallbacks"]),
                "exclusive_tasks": int((batch_plan & _abi.PLAN_EXCLUSIVE != 0).sum()),
                "edge_warp_kernel_tasks": int((batch_plan & _abi.PLAN_EDGE_WARP != 0).sum()),
                "per_rank": {"fields": ["device_ms_per_step", "cells_per_step", "segmented_tasks", "segment_fallbacks", "exclusive_tasks", "tasks"],
                             "values": per_rank},
                "timing": "sum over steps of CUDA-event time on the library's launch stream, max over ranks"}
        print(json.dumps(line), flush=True)
    al.close()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and parity is not None and not parity["bit_exact"]:
        sys.stderr.write("bench.py: PARITY FAILED: %s\n" % json.dumps(parity))
        sys.exit(3)


if __name__ == "__main__":
    main()
