#!/bin/bash
# round 2: 2-GPU weak-scaling check (per-rank device time in the bench line)
out=gpurun_out; mkdir -p $out; tag=${1:-r2n2}
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline > $out/bench_${tag}.jsonl 2> $out/${tag}.err; echo "bench N=2 rc $?"; tail -3 $out/${tag}.err
python -c "
import json;d=json.loads([l for l in open('$out/bench_${tag}.jsonl') if l.startswith('{')][-1]);print({k:d[k] for k in ('value','ms_per_step','n_gpus','segmented_tasks','segment_fallbacks')});print(d['e2e']);print(d['parity']);print(d['per_rank'])"
