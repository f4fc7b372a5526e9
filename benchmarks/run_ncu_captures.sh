#!/bin/bash
# Full ncu captures of the three kernels that head the step's time table (one GPU; results in gpurun_out/ncu_r2/).
set -u
out=gpurun_out/ncu_r2; mkdir -p $out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 170 $NCU -k regex:conv_gemm_halo_kernel -s 2 -c 1 -o $out/dgrad3_halo python benchmarks/run_one_conv.py conv.3 dgrad default > $out/dgrad3.log 2>&1
timeout 170 $NCU -k regex:conv_gemm_persistent_kernel -s 2 -c 1 -o $out/dgrad10_pair python benchmarks/run_one_conv.py conv.10 dgrad default > $out/dgrad10.log 2>&1
timeout 170 $NCU -k regex:conv_gemm_persistent_kernel -s 2 -c 1 -o $out/fwd10_pair python benchmarks/run_one_conv.py conv.10 fwd default > $out/fwd10.log 2>&1
# whole step: time + DRAM bytes per kernel (eager fused step, every kernel once)
timeout 250 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --profile-from-start off --csv --log-file $out/step_bytes.csv python benchmarks/profile_step.py > $out/step.log 2>&1
ls -la $out
