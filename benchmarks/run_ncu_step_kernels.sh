#!/bin/bash
# Full ncu captures of kernels as the fused step launches them (eager step, profiler range around one step).
#   bash benchmarks/run_ncu_step_kernels.sh "<kernel regex>" <name> ["<kernel regex>" <name> ...]
set -u
out=gpurun_out/ncu_r2; mkdir -p $out
while [ $# -ge 2 ]; do
  timeout 200 ncu --set full --clock-control none --import-source on -f --profile-from-start off -k "regex:$1" -c 1 \
    -o $out/$2 python benchmarks/profile_step.py > $out/$2.log 2>&1
  shift 2
done
ls -la $out | tail -8
