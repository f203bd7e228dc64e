#!/bin/bash
# Round-end evidence run on one B200 (gpurun -- 'bash scripts/round_end.sh <tag>'): GPU tests, bench lines, all configs,
# segmented-vs-whole probes, then the two ncu passes (launch list, one full capture of the dominant kernel).
tag=${1:-r1f}; out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/${tag}_gputests.log 2>&1; echo "gpu tests rc $?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc $?"
timeout 600 python bench.py > $out/bench_${tag}_n1.jsonl 2> $out/${tag}_n1.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference > $out/bench_${tag}_reference.jsonl 2> $out/${tag}_ref.err; echo "ref rc $?"
timeout 600 python scripts/run_configs.py 0.25 > $out/configs_${tag}.log 2>&1; echo "configs rc $?"
timeout 600 python scripts/run_configs.py 1.0 4 cfg1,cfg3,cfg4 > $out/configs_full_${tag}.log 2>&1; echo "full configs rc $?"
timeout 300 python bench.py --workload cfg3 > $out/bench_${tag}_cfg3.jsonl 2> $out/${tag}_cfg3.err; echo "cfg3 rc $?"
( timeout 200 python scripts/segcheck.py asm5 200000 3001 2; timeout 200 python scripts/segcheck.py asm10 1000000 3001 2;
  timeout 200 python scripts/segcheck.py hifiasm 150000 500 4 ) > $out/segcheck_${tag}.log 2>&1; echo "segcheck rc $?"
timeout 300 python scripts/kbench.py asm5 20000 3001 592 > $out/kbench_${tag}.log 2>&1; echo "kbench rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_${tag}.csv \
  python bench.py --regions 600 --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_launch.log 2>&1; echo "ncu launches rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fsv_fill_dpx -s 1 -c 1 -f -o $out/prof_dpx_${tag} \
  python scripts/kbench.py asm5 20000 3001 592 > $out/${tag}_ncu.log 2>&1; echo "ncu full rc $?"
tail -3 $out/${tag}_gputests.log; cat $out/${tag}_smoke.log | tail -2; cat $out/configs_${tag}.log; cat $out/configs_full_${tag}.log | tail -8; cat $out/segcheck_${tag}.log | grep -v "^dump"
