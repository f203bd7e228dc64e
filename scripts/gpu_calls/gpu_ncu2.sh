#!/bin/bash
# ncu --set full of the edge-warp kernel on a probe: gpu_ncu2.sh <tag> <kbench args...>
out=gpurun_out; mkdir -p $out; tag=$1; shift
timeout 100 python scripts/kbench.py "$@" | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fsv_fill_ -s 1 -c 1 -f -o $out/prof_${tag} python scripts/kbench.py "$@" > $out/${tag}_ncu.log 2>&1; echo "ncu full rc $?"
