#!/bin/bash
# usage: gpu_ncu.sh <op> <kernel regex> [<op> <kernel regex> ...]  -- plain run first, then ncu --set full with source
mkdir -p gpurun_out
while [ $# -ge 2 ]; do
  op=$1; k=$2; shift 2
  timeout 300 python scripts/prof.py $op --time > gpurun_out/prof_$op.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/ncu_$op python scripts/prof.py $op > gpurun_out/ncu_$op.log 2>&1
  tail -n 3 gpurun_out/prof_$op.log gpurun_out/ncu_$op.log
done
