#!/bin/bash
# round 2, GPU call N: default cold-start lead of a segment 300 % of the band: GPU tests, every-task parity of all configs, bench
out=gpurun_out; mkdir -p $out; tag=${1:-r2n}
timeout 400 python -m pytest tests -m gpu -x -q -p timeout --timeout 150 > $out/${tag}_gputests.log 2>&1; echo "gpu tests rc $?"; tail -4 $out/${tag}_gputests.log
( for c in cfg1 cfg3 cfg4 long1m; do timeout 300 python scripts/parity_full.py gpu $c; done ) > $out/parity_full_${tag}.log 2>&1; echo "parity_full rc $?"; grep -E "MISMATCH|BIT-EXACT|task " $out/parity_full_${tag}.log
timeout 600 python scripts/run_configs.py 1.0 4 cfg1,cfg3,cfg4 > $out/configs_full_${tag}.log 2>&1; echo "full configs rc $?"; cut -c1-200 $out/configs_full_${tag}.log
timeout 600 python bench.py --steps 5 --warmup 3 > $out/bench_${tag}_n1.jsonl 2> $out/${tag}_n1.err; echo "bench rc $?"; cut -c1-200 $out/bench_${tag}_n1.jsonl
timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 > $out/bench_${tag}_cfg3.jsonl 2> $out/${tag}_cfg3.err; echo "cfg3 rc $?"; cut -c1-200 $out/bench_${tag}_cfg3.jsonl
