#!/bin/bash
# Round-end evidence run on one B200 (gpurun -- 'bash scripts/round_end.sh <tag>'): GPU tests, bench lines, all configs at full size with
# every task compared with the oracle digests, kernel probes, then the ncu passes (launch list, DRAM traffic of one step, one full
# capture of the dominant kernel).  Everything lands in gpurun_out/; what is to be judged is copied into profiles/ afterwards.
tag=${1:-r2}; out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1; nproc >> $out/${tag}_smi.txt
timeout 600 python -m pytest tests -m gpu -q -p timeout --timeout 200 > $out/${tag}_gputests.log 2>&1; echo "gpu tests rc $?"; tail -3 $out/${tag}_gputests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc $?"; tail -1 $out/${tag}_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > $out/bench_${tag}_n1.jsonl 2> $out/${tag}_n1.err; echo "bench rc $?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_${tag}_reference.jsonl 2> $out/${tag}_ref.err; echo "ref rc $?"
timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 > $out/bench_${tag}_cfg3.jsonl 2> $out/${tag}_cfg3.err; echo "cfg3 rc $?"
timeout 300 python bench.py --workload cfg3 --impl reference --steps 2 --warmup 1 > $out/bench_${tag}_cfg3_reference.jsonl 2>> $out/${tag}_cfg3.err; echo "cfg3 ref rc $?"
( for c in cfg1 cfg3 cfg4 long1m; do timeout 300 python scripts/parity_full.py gpu $c; done ) > $out/parity_full_${tag}.log 2>&1; echo "parity_full rc $?"; grep -E "MISMATCH|BIT-EXACT" $out/parity_full_${tag}.log
timeout 600 python scripts/run_configs.py 1.0 4 cfg1,cfg3,cfg4 > $out/configs_full_${tag}.log 2>&1; echo "full configs rc $?"
( timeout 100 python scripts/kbench.py asm5 20000 3001 592; timeout 100 python scripts/kbench.py hifiasm 20000 500 1184; timeout 100 python scripts/kbench.py map-hifi 18000 751 1184; timeout 100 python scripts/kbench.py asm5 20000 3001 592 0x2; timeout 100 python scripts/kbench.py hifiasm 20000 500 1184 0x19 ) > $out/kbench_${tag}.log 2>&1; echo "kbench rc $?"
timeout 200 python scripts/chainbench.py 200 > $out/chainbench_${tag}.log 2>&1; echo "chainbench rc $?"; cat $out/chainbench_${tag}.log
timeout 200 python scripts/edbench.py 4000 > $out/edbench_${tag}.log 2>&1; echo "edbench rc $?"; cat $out/edbench_${tag}.log
FSV_TRACE=1 timeout 200 python scripts/shardprobe.py 1 > $out/shardprobe_${tag}.log 2>&1; echo "shardprobe rc $?"
# ---- ncu (numbers printed under ncu are never bench values)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_${tag}.csv \
  python bench.py --regions 600 --steps 2 --warmup 1 --no-cpu-baseline --no-parity > $out/${tag}_ncu_launch.log 2>&1; echo "ncu launches rc $?"
timeout 900 ncu --replay-mode application --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fsv_fill --csv --log-file $out/ncu_traffic_${tag}.csv \
  python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-parity > $out/${tag}_ncu_traffic.log 2>&1; echo "ncu traffic rc $?"
python scripts/ncu_traffic.py $out/ncu_traffic_${tag}.csv $out/bench_${tag}_n1.jsonl $out/ncu_traffic_${tag}.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fsv_fill_dpx -s 1 -c 1 -f -o $out/prof_dpx_${tag} \
  python scripts/kbench.py asm5 20000 3001 592 > $out/${tag}_ncu.log 2>&1; echo "ncu full rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fsv_fill_ew -s 1 -c 1 -f -o $out/prof_ew_${tag} \
  python scripts/kbench.py hifiasm 20000 500 1184 > $out/${tag}_ncu_ew.log 2>&1; echo "ncu full (edge-warp kernel) rc $?"
cut -c1-300 $out/bench_${tag}_n1.jsonl; cat $out/configs_full_${tag}.log | cut -c1-260; grep GCUPS $out/kbench_${tag}.log | awk 'NR%3==0'
