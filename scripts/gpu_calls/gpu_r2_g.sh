#!/bin/bash
# round 2, GPU call G: kernel iteration check - GPU tests, kernel probes, every-task parity of cfg1 / cfg3 / cfg4 / long1m
out=gpurun_out; mkdir -p $out; tag=${1:-r2g}
timeout 400 python -m pytest tests -m gpu -x -q -p timeout --timeout 150 > $out/${tag}_gputests.log 2>&1; echo "gpu tests rc $?"; tail -4 $out/${tag}_gputests.log
( timeout 60 python scripts/kbench.py asm5 20000 3001 592 | tail -1
  timeout 60 python scripts/kbench.py asm5 20000 3001 592 | tail -1
  timeout 60 python scripts/kbench.py hifiasm 20000 500 1184 | tail -1
  timeout 60 python scripts/kbench.py map-hifi 18000 751 1184 | tail -1
  timeout 60 python scripts/kbench.py asm5 20000 3001 592 0x2 | tail -1
  timeout 60 python scripts/kbench.py hifiasm 20000 500 1184 0x19 | tail -1 ) > $out/kbench_${tag}.log 2>&1; cat $out/kbench_${tag}.log
( for c in ${2:-cfg1 cfg3 cfg4 long1m}; do timeout 300 python scripts/parity_full.py gpu $c; done ) > $out/parity_full_${tag}.log 2>&1; echo "parity_full rc $?"; grep -E "MISMATCH|BIT-EXACT" $out/parity_full_${tag}.log
