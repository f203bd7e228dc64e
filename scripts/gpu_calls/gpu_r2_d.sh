#!/bin/bash
# round 2, GPU call D: fixed 2-instruction subtraction kernel (parity + speed), LPT-simulated auto segmentation, sweeps
out=gpurun_out; mkdir -p $out; tag=${1:-r2d}
timeout 500 python -m pytest tests -m gpu -x -q --durations=5 -p timeout --timeout 150 > $out/${tag}_gputests.log 2>&1; echo "gpu tests rc $?"; tail -12 $out/${tag}_gputests.log
( timeout 60 python scripts/kbench.py asm5 20000 3001 592; timeout 60 python scripts/kbench.py hifiasm 20000 500 1184; timeout 60 python scripts/kbench.py map-hifi 18000 751 1184 ) > $out/kbench_${tag}.log 2>&1; echo "kbench rc $?"; grep GCUPS $out/kbench_${tag}.log | awk 'NR%3==0'
FSV_TRACE=1 timeout 300 python scripts/segsweep.py cfg2 "segment_auto_pct=50;segment_auto_pct=70;segment_auto_pct=100;segment_auto_pct=70,segment_warm_pct=300;segment_min_diags=0" > $out/${tag}_sweep_cfg2.log 2>&1; echo "sweep cfg2 rc $?"; grep -v "^\[fsv\] \(create\|run\|fetch\|destroy\)" $out/${tag}_sweep_cfg2.log | tail -20
FSV_TRACE=1 timeout 120 python scripts/segsweep.py cfg3 "segment_auto_pct=70;segment_auto_pct=100" > $out/${tag}_sweep_cfg3.log 2>&1; echo "sweep cfg3 rc $?"; grep -v "^\[fsv\] \(create\|run\|fetch\|destroy\)" $out/${tag}_sweep_cfg3.log | tail
FSV_TRACE=1 timeout 150 python scripts/segsweep.py cfg4 "segment_auto_pct=70;segment_auto_pct=100" > $out/${tag}_sweep_cfg4.log 2>&1; echo "sweep cfg4 rc $?"; grep -v "^\[fsv\] \(create\|run\|fetch\|destroy\)" $out/${tag}_sweep_cfg4.log | tail
FSV_TRACE=1 timeout 90 python scripts/segsweep.py cfg1 "segment_auto_pct=70;segment_auto_pct=70,segment_warm_pct=600" > $out/${tag}_sweep_cfg1.log 2>&1; echo "sweep cfg1 rc $?"; grep -v "^\[fsv\] \(create\|run\|fetch\|destroy\)" $out/${tag}_sweep_cfg1.log | tail
