#!/bin/bash
# one ncu --set full capture of the dominant kernel on the uniform probe (after the same command ran without ncu)
out=gpurun_out; mkdir -p $out; tag=${1:-x}
timeout 100 python scripts/kbench.py asm5 20000 3001 592 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fsv_fill_dpx -s 1 -c 1 -f -o $out/prof_dpx_${tag} python scripts/kbench.py asm5 20000 3001 592 > $out/${tag}_ncu.log 2>&1; echo "ncu full rc $?"
