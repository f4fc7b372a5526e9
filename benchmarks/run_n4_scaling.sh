#!/bin/bash
# Why does the step cost +150 us at 4 / 8 GPUs but +29 us at 2?  Same box, 4 GPUs, device-timed loop only.
OUT=gpurun_out/n4exp
mkdir -p $OUT
export NCCL_DEBUG=WARN
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node"
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2> $OUT/$name.err; echo "rc=$? $(tail -c 400 $OUT/$name.log | tr '\n' ' ' | cut -c1-400)"; }
port=29700
for cfg in "A=1" "DDPT_NO_MULTICAST=1" "DDPT_COMM_BLOCKS=32" "DDPT_COMM_BLOCKS=16" "DDPT_ALLREDUCE_ALGO=twoshot" "DDPT_COMM_WIRE=bf16"; do
  port=$((port+1))
  run "bench_$cfg" 120 env $cfg $TR 4 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 4 --steps 200 --warmup 10 --skip_e2e
done
run bench_n1 100 python bench.py --steps 200 --warmup 10 --skip_e2e
run tests 200 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 150 -s -k "three_gpus or buffer_and_parameter"
run resnet_ours 150 $TR 4 --master-addr 127.0.0.1 --master-port 29721 benchmarks/bench_resnet50.py --impl ours
run resnet_ours_cudnn 150 $TR 4 --master-addr 127.0.0.1 --master-port 29722 benchmarks/bench_resnet50.py --impl ours --no_tc
run resnet_torch 150 $TR 4 --master-addr 127.0.0.1 --master-port 29723 benchmarks/bench_resnet50.py --impl torch
echo done
