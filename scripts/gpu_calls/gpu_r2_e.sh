#!/bin/bash
# round 2, GPU call E: A/B of the subtraction formula on one box, planner fix (all configs), bench line
out=gpurun_out; mkdir -p $out; tag=${1:-r2e}
timeout 400 python -m pytest tests -m gpu -x -q -p timeout --timeout 150 > $out/${tag}_gputests.log 2>&1; echo "gpu tests rc $?"; tail -4 $out/${tag}_gputests.log
for rep in 1 2; do
  for lib in libfocalsv_cuda.so libfsv_oldsub.so; do
    echo "== $lib (rep $rep)"; FSV_LIB_PATH=$PWD/focalsv_b200/$lib timeout 60 python scripts/kbench.py asm5 20000 3001 592 | tail -1
    FSV_LIB_PATH=$PWD/focalsv_b200/$lib timeout 60 python scripts/kbench.py hifiasm 20000 500 1184 | tail -1
  done
done > $out/kbench_ab_${tag}.log 2>&1; cat $out/kbench_ab_${tag}.log
FSV_TRACE=1 timeout 200 python scripts/segsweep.py cfg2 "segment_auto_pct=70;segment_auto_pct=60;segment_auto_pct=80" > $out/${tag}_sweep_cfg2.log 2>&1; echo "sweep cfg2 rc $?"; grep -v "^\[fsv\] \(create\|run\|fetch\|destroy\)" $out/${tag}_sweep_cfg2.log | tail -8
FSV_TRACE=1 timeout 200 python scripts/segsweep.py cfg2.r1w2 "segment_auto_pct=70" > $out/${tag}_sweep_cfg2r1w2.log 2>&1; echo "sweep cfg2.r1w2 rc $?"; grep -v "^\[fsv\] \(create\|run\|fetch\|destroy\)" $out/${tag}_sweep_cfg2r1w2.log | tail -3
FSV_TRACE=1 timeout 120 python scripts/segsweep.py cfg3 "segment_auto_pct=70;segment_auto_pct=50" > $out/${tag}_sweep_cfg3.log 2>&1; echo "sweep cfg3 rc $?"; grep -v "^\[fsv\] \(create\|run\|fetch\|destroy\)" $out/${tag}_sweep_cfg3.log | tail
FSV_TRACE=1 timeout 150 python scripts/segsweep.py cfg4 "segment_auto_pct=70;segment_auto_pct=50" > $out/${tag}_sweep_cfg4.log 2>&1; echo "sweep cfg4 rc $?"; grep -v "^\[fsv\] \(create\|run\|fetch\|destroy\)" $out/${tag}_sweep_cfg4.log | tail
FSV_TRACE=1 timeout 90 python scripts/segsweep.py cfg1 "segment_auto_pct=70" > $out/${tag}_sweep_cfg1.log 2>&1; echo "sweep cfg1 rc $?"; grep -v "^\[fsv\] \(create\|run\|fetch\|destroy\)" $out/${tag}_sweep_cfg1.log | tail
timeout 400 python bench.py --steps 3 --warmup 3 > $out/bench_${tag}_n1.jsonl 2> $out/${tag}_n1.err; echo "bench rc $?"; cut -c1-250 $out/bench_${tag}_n1.jsonl; python -c "
import json;d=json.load(open('$out/bench_${tag}_n1.jsonl'));print({k:d[k] for k in ('value','ms_per_step','segmented_tasks','segment_fallbacks','exclusive_tasks')});print(d['e2e']);print(d['parity']);print(d['cpu_baseline']);print(d['roofline']['frac'])"
