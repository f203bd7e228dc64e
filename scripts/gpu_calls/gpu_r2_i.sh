#!/bin/bash
# round 2, GPU call I: edge-warp kernel - parity after the key fix; per-antidiagonal time of one CTA per SM (148 tasks), old kernel vs edge-warp kernel with 255 registers
out=gpurun_out; mkdir -p $out; tag=${1:-r2i}
timeout 400 python -m pytest tests -m gpu -x -q -p timeout --timeout 150 > $out/${tag}_gputests.log 2>&1; echo "gpu tests rc $?"; tail -4 $out/${tag}_gputests.log
( echo "== ew kernel (2 CTAs/SM, 128 regs)"; timeout 60 python scripts/kbench.py asm5 20000 3001 592 | tail -1
  echo "== old kernel"; timeout 60 python scripts/kbench.py asm5 20000 3001 592 0 0 ew_kernel=0 | tail -1
  echo "== 148 tasks: old kernel, old kernel exclusive, ew 128 regs, ew 255 regs (1 CTA/SM)"
  timeout 60 python scripts/kbench.py asm5 20000 3001 148 0 0 ew_kernel=0 | tail -1
  timeout 60 python scripts/kbench.py asm5 20000 3001 148 0 0 ew_kernel=0 force_excl=1 | tail -1
  timeout 60 python scripts/kbench.py asm5 20000 3001 148 | tail -1
  FSV_LIB_PATH=$PWD/focalsv_b200/libfsv_ewocc1.so timeout 60 python scripts/kbench.py asm5 20000 3001 148 | tail -1
  echo "== narrow bands: ew / old"
  timeout 60 python scripts/kbench.py hifiasm 20000 500 1184 | tail -1
  timeout 60 python scripts/kbench.py hifiasm 20000 500 1184 0 0 ew_kernel=0 | tail -1
  timeout 60 python scripts/kbench.py map-hifi 18000 751 1184 | tail -1
  timeout 60 python scripts/kbench.py map-hifi 18000 751 1184 0 0 ew_kernel=0 | tail -1 ) > $out/kbench_${tag}.log 2>&1; cat $out/kbench_${tag}.log
( for c in ${2:-cfg1 cfg3 cfg4 long1m}; do timeout 300 python scripts/parity_full.py gpu $c; done ) > $out/parity_full_${tag}.log 2>&1; echo "parity_full rc $?"; grep -E "MISMATCH|BIT-EXACT|task " $out/parity_full_${tag}.log
