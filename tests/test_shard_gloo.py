"""The N > 1 path of bench.py on CPU: world_size-2 gloo, LPT sharding of the region batch with no
data-path collective (regions are independent); only a barrier and a max/sum over ranks."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
import bench
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
group, n_regions = bench.build_shard(rank, world, 40)
t = group.tasks
est = float(((t["qlen"].astype(np.int64) + t["tlen"] - 1) * np.minimum(np.minimum(t["qlen"], t["tlen"]), t["w"] + 1)).sum())
v = torch.tensor([est, float(n_regions), float(len(t))], dtype=torch.float64)
allv = [torch.zeros_like(v) for _ in range(world)]
dist.all_gather(allv, v)
mx = v.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
dist.barrier()
if rank == 0:
    print(json.dumps({"per_rank": [x.tolist() for x in allv], "max": mx.tolist()}))
dist.destroy_process_group()
''' % ROOT


def test_two_rank_sharding_is_disjoint_complete_and_balanced(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    per = np.array(d["per_rank"])
    assert per[:, 1].sum() == 80                 # 2 ranks x 40 regions, every region on exactly one rank
    assert (per[:, 2] == 2 * per[:, 1]).all()    # two haplotype contigs per region
    assert per[:, 0].max() / per[:, 0].mean() < 1.25   # length-balanced bins


def test_single_rank_shard_is_the_whole_config():
    sys.path.insert(0, ROOT)
    import bench
    g, n = bench.build_shard(0, 1, 30)
    assert n == 30 and len(g.tasks) == 60
    assert int(g.tasks["w"][0]) == 3001 and int(g.tasks["zdrop"][0]) == 200 and g.scoring.q2 == 81
