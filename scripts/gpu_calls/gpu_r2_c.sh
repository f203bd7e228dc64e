#!/bin/bash
# round 2, GPU call C: 2-instruction subtraction kernel (parity + speed), segmentation sweeps
out=gpurun_out; mkdir -p $out; tag=${1:-r2c}
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > $out/${tag}_gputests.log 2>&1; echo "gpu tests rc $?"; tail -12 $out/${tag}_gputests.log
( timeout 200 python scripts/kbench.py asm5 20000 3001 592; timeout 200 python scripts/kbench.py hifiasm 20000 500 1184; timeout 200 python scripts/kbench.py map-hifi 18000 751 1184 ) > $out/kbench_${tag}.log 2>&1; echo "kbench rc $?"; grep GCUPS $out/kbench_${tag}.log | awk 'NR%3==0'
timeout 600 python scripts/segsweep.py cfg2 "segment_auto_pct=35;segment_auto_pct=50;segment_auto_pct=70;segment_auto_pct=100;segment_auto_pct=50,segment_warm_pct=300;segment_auto_pct=50,segment_warm_pct=250;segment_auto_pct=50,segment_rows=106496;segment_min_diags=0" > $out/${tag}_sweep_cfg2.log 2>&1; echo "sweep cfg2 rc $?"; cat $out/${tag}_sweep_cfg2.log
timeout 300 python scripts/segsweep.py cfg3 "segment_auto_pct=35;segment_auto_pct=70;segment_auto_pct=100;segment_auto_pct=200;segment_min_diags=0" > $out/${tag}_sweep_cfg3.log 2>&1; echo "sweep cfg3 rc $?"; cat $out/${tag}_sweep_cfg3.log
timeout 300 python scripts/segsweep.py cfg4 "segment_auto_pct=35;segment_auto_pct=70;segment_auto_pct=100" > $out/${tag}_sweep_cfg4.log 2>&1; echo "sweep cfg4 rc $?"; cat $out/${tag}_sweep_cfg4.log
timeout 300 python scripts/segsweep.py cfg1 "segment_auto_pct=35;segment_auto_pct=35,segment_warm_pct=300;segment_auto_pct=35,segment_warm_pct=600" > $out/${tag}_sweep_cfg1.log 2>&1; echo "sweep cfg1 rc $?"; cat $out/${tag}_sweep_cfg1.log
