"""Row f2 on the GPU: cfg1's contig pairs (and n regions of cfg2's shape) aligned as ONE global band-3001 task per pair
vs decomposed by fsv_chain_pieces into small fills (hook.realign_regions_chained).  usage: chainbench.py [n_regions]"""
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np
from focalsv_b200 import api, hook, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
al = api.Aligner(0)
for name, groups in (("cfg1.contigs", synth.config1(n_reads=8)[:1]), ("cfg2 x %d regions" % n, synth.config2(n_regions=n, max_region=400000)[:1])):
    g = groups[0]
    windows = [("chr21", 0, g.tarena[int(t["t_off"]):int(t["t_off"]) + int(t["tlen"])]) for t in g.tasks]
    contigs = [("c%d" % i, g.qarena[int(t["q_off"]):int(t["q_off"]) + int(t["qlen"])]) for i, t in enumerate(g.tasks)]
    for rep in range(2):
        t0 = time.perf_counter(); whole = hook.realign_regions(al, windows, contigs, preset="asm5", bw=2000); t_whole = time.perf_counter() - t0
        k0 = al.stats()["total_ms"]
        t0 = time.perf_counter(); ch = hook.realign_regions_chained(al, windows, contigs, preset="asm5", bw=2000); t_ch = time.perf_counter() - t0
        k1 = al.stats()["total_ms"]
    sw = hook.signatures(whole); sc = hook.signatures(ch)
    same = sorted((s.qname, s.svtype, s.svlen) for s in sw if s.svlen >= 50) == sorted((s.qname, s.svtype, s.svlen) for s in sc if s.svlen >= 50)
    nz = sum(1 for r in whole if r.zdropped)
    print("%-22s pairs %4d  whole: %8.1f ms wall (kernels %.1f ms, %d z-dropped)  chained: %8.1f ms wall (kernels %.1f ms)  SV signatures >= 50 bp equal: %s (%d vs %d)" % (
        name, len(windows), t_whole * 1e3, k0, nz, t_ch * 1e3, k1, same, len(sw), len(sc)), flush=True)
