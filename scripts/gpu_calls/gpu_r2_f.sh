#!/bin/bash
# round 2, GPU call F: complemented-z formulation of dpx_cells (2 instructions fewer per word, complements on the FMA pipe), one-LOP3 traceback byte
out=gpurun_out; mkdir -p $out; tag=${1:-r2f}
timeout 400 python -m pytest tests -m gpu -x -q -p timeout --timeout 150 > $out/${tag}_gputests.log 2>&1; echo "gpu tests rc $?"; tail -4 $out/${tag}_gputests.log
for rep in 1 2; do
  for lib in libfocalsv_cuda.so libfsv_notlop3.so; do
    echo "== $lib (rep $rep)"; FSV_LIB_PATH=$PWD/focalsv_b200/$lib timeout 60 python scripts/kbench.py asm5 20000 3001 592 | tail -1
    FSV_LIB_PATH=$PWD/focalsv_b200/$lib timeout 60 python scripts/kbench.py hifiasm 20000 500 1184 | tail -1
  done
done > $out/kbench_ab_${tag}.log 2>&1; cat $out/kbench_ab_${tag}.log
timeout 60 python scripts/kbench.py map-hifi 18000 751 1184 | tail -1
timeout 60 python scripts/kbench.py asm5 20000 3001 592 0x2 | tail -1
( for c in cfg1 cfg3; do timeout 300 python scripts/parity_full.py gpu $c; done ) > $out/parity_full_${tag}.log 2>&1; echo "parity_full rc $?"; grep -E "MISMATCH|BIT-EXACT" $out/parity_full_${tag}.log
