#!/bin/bash
# round 2: N-GPU weak-scaling check (per-rank device time in the bench line); usage: gpu_r2_n8.sh <tag> <N>
out=gpurun_out; mkdir -p $out; tag=${1:-r2n8}; N=${2:-8}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline > $out/bench_${tag}.jsonl 2> $out/${tag}.err; echo "bench N=$N rc $?"; tail -3 $out/${tag}.err
python -c "
import json;d=json.loads([l for l in open('$out/bench_${tag}.jsonl') if l.startswith('{')][-1]);print({k:d[k] for k in ('value','ms_per_step','n_gpus','segmented_tasks','segment_fallbacks')});print(d['e2e']['value']);print(d['parity']['bit_exact']);print([round(v[0]) for v in d['per_rank']['values']])"
