#!/bin/bash
# compute-sanitizer over the kernel tests: memcheck on the GEMM / elementwise / optimizer kernels,
# racecheck + synccheck on the shared-memory users (BatchNorm passes, head, conv.0, fused optimizer).
#   gpurun -- bash benchmarks/run_sanitize.sh        -> gpurun_out/sanitize/*.log, summaries copied to profiles/sanitize/
OUT=gpurun_out/sanitize
mkdir -p $OUT
SAN="compute-sanitizer --error-exitcode 1 --print-limit 20"
K1='(conv3x3_fwd_and_stats or conv3x3_dgrad or conv3x3_wgrad or linear_fwd or linear_dgrad or linear_wgrad or fused_lookahead or multi_tensor or scaler) and not mxfp8 and not (1024 and 8192)'
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' $OUT/$name.log | tail -3 | tr '\n' ' ')"; }
run memcheck_kernels 420 $SAN --tool memcheck python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "$K1"
run memcheck_toynet 300 $SAN --tool memcheck python -m pytest tests/test_gpu_toynet.py -m gpu -q -x -k "forward_train_and_eval or backward_matches or deterministic"
run racecheck_toynet 300 $SAN --tool racecheck python -m pytest tests/test_gpu_toynet.py -m gpu -q -x -k "backward_matches"
run synccheck_toynet 300 $SAN --tool synccheck python -m pytest tests/test_gpu_toynet.py -m gpu -q -x -k "backward_matches"
run memcheck_comm 240 $SAN --tool memcheck python -m pytest tests/test_gpu_comm.py -m gpu -q -x -k "world2 or 2- or bit_identical or symm_broadcast"
for f in $OUT/*.log; do echo "--- $f"; grep -E "ERROR SUMMARY|passed|failed|error" $f | tail -4; done > $OUT/SUMMARY.txt
cat $OUT/SUMMARY.txt
