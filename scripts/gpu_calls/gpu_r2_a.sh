#!/bin/bash
# round 2, GPU call A: full GPU test suite (incl. all-task parity of the timed configs), shard straggler probe, bench line, kernel probes
out=gpurun_out; mkdir -p $out; tag=${1:-r2a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $out/${tag}_smi.txt 2>&1; nproc >> $out/${tag}_smi.txt; free -g >> $out/${tag}_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q --durations=15 > $out/${tag}_gputests.log 2>&1; echo "gpu tests rc $?"; tail -25 $out/${tag}_gputests.log
FSV_TRACE=1 timeout 400 python scripts/shardprobe.py 2 > $out/${tag}_shard_w2.log 2>&1; echo "shardprobe w2 rc $?"
FSV_TRACE=1 timeout 300 python scripts/shardprobe.py 1 > $out/${tag}_shard_w1.log 2>&1; echo "shardprobe w1 rc $?"
grep -v "^\[fsv\] \(create\|run\|fetch\|destroy\)" $out/${tag}_shard_w2.log | tail -40; grep -v "^\[fsv\] \(create\|run\|fetch\|destroy\)" $out/${tag}_shard_w1.log | tail -20
timeout 600 python bench.py --steps 3 --warmup 3 > $out/bench_${tag}_n1.jsonl 2> $out/${tag}_n1.err; echo "bench rc $?"; cut -c1-600 $out/bench_${tag}_n1.jsonl; tail -3 $out/${tag}_n1.err
( timeout 200 python scripts/kbench.py asm5 20000 3001 592; timeout 200 python scripts/kbench.py hifiasm 20000 500 1184; timeout 200 python scripts/kbench.py map-hifi 18000 751 1184 ) > $out/kbench_${tag}.log 2>&1; echo "kbench rc $?"; grep GCUPS $out/kbench_${tag}.log | awk 'NR%3==0'
