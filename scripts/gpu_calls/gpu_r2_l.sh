#!/bin/bash
# round 2, GPU call L: traceback pages kept across a CTA's tasks - GPU tests, small-task batches (chained path, all configs), bench
out=gpurun_out; mkdir -p $out; tag=${1:-r2l}
timeout 400 python -m pytest tests -m gpu -x -q -p timeout --timeout 150 > $out/${tag}_gputests.log 2>&1; echo "gpu tests rc $?"; tail -4 $out/${tag}_gputests.log
timeout 200 python scripts/chainbench.py 200 > $out/chainbench_${tag}.log 2>&1; echo "chainbench rc $?"; cat $out/chainbench_${tag}.log
( for c in ${2:-cfg1 cfg3 cfg4 long1m}; do timeout 300 python scripts/parity_full.py gpu $c; done ) > $out/parity_full_${tag}.log 2>&1; echo "parity_full rc $?"; grep -E "MISMATCH|BIT-EXACT|task " $out/parity_full_${tag}.log
timeout 600 python scripts/run_configs.py 1.0 4 cfg1,cfg3,cfg4 > $out/configs_full_${tag}.log 2>&1; echo "full configs rc $?"; cut -c1-200 $out/configs_full_${tag}.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_${tag}_n1.jsonl 2> $out/${tag}_n1.err; echo "bench rc $?"; cut -c1-200 $out/bench_${tag}_n1.jsonl
timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_${tag}_cfg3.jsonl 2> $out/${tag}_cfg3.err; echo "cfg3 rc $?"; cut -c1-200 $out/bench_${tag}_cfg3.jsonl
