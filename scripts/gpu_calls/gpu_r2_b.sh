#!/bin/bash
# round 2, GPU call B: segment rework (dynamic pages + repair): tests, shard probes, bench
out=gpurun_out; mkdir -p $out; tag=${1:-r2b}
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > $out/${tag}_gputests.log 2>&1; echo "gpu tests rc $?"; tail -22 $out/${tag}_gputests.log
FSV_TRACE=1 timeout 300 python scripts/shardprobe.py 1 > $out/${tag}_shard_w1.log 2>&1; echo "shardprobe w1 rc $?"
FSV_TRACE=1 timeout 400 python scripts/shardprobe.py 2 > $out/${tag}_shard_w2.log 2>&1; echo "shardprobe w2 rc $?"
grep -v "^\[fsv\] \(create\|run\|fetch\|destroy\)" $out/${tag}_shard_w1.log | tail -16; grep -v "^\[fsv\] \(create\|run\|fetch\|destroy\)" $out/${tag}_shard_w2.log | tail -32
timeout 600 python scripts/run_configs.py 1.0 4 cfg1,cfg3,cfg4 > $out/configs_full_${tag}.log 2>&1; echo "full configs rc $?"; cat $out/configs_full_${tag}.log
