#!/bin/bash
# One 8-GPU session (gpurun --gpus 8 is charged 8x): correctness of the NVLS paths first, then every 8-GPU / 4-GPU number.
#   gpurun --gpus 8 -- bash benchmarks/run_n8.sh
OUT=gpurun_out/n8r2
mkdir -p $OUT
export NCCL_DEBUG=WARN
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2> $OUT/$name.err; echo "rc=$? $(tail -c 300 $OUT/$name.log | tr '\n' ' ' | cut -c1-300)"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node"
run tests 300 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 200 -s -k "(matches_nccl and torch) or engine_world2 or three_gpus or buffer_and_parameter"
run sweep_n8 150 $TR 8 --master-addr 127.0.0.1 --master-port 29620 benchmarks/allreduce_sweep.py --iters 10 --out $OUT/allreduce_sweep_n8.json
run sweep_n4 150 $TR 4 --master-addr 127.0.0.1 --master-port 29624 benchmarks/allreduce_sweep.py --iters 10 --out $OUT/allreduce_sweep_n4.json
run bench_n1 100 python bench.py --steps 200 --warmup 10
run bench_n8 150 $TR 8 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 8 --steps 200 --warmup 10 --census $OUT/census_n8.json
run bench_n8_bf16 150 $TR 8 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 8 --steps 200 --warmup 10 --comm_wire bf16
run bench_n4 150 $TR 4 --master-addr 127.0.0.1 --master-port 29603 bench.py --gpus 4 --steps 200 --warmup 10
run bench_n2 150 $TR 2 --master-addr 127.0.0.1 --master-port 29604 bench.py --gpus 2 --steps 200 --warmup 10
run bench_ref_n8 200 $TR 8 --master-addr 127.0.0.1 --master-port 29605 bench.py --gpus 8 --impl reference --steps 40 --warmup 5
run resnet_ours 150 $TR 8 --master-addr 127.0.0.1 --master-port 29621 benchmarks/bench_resnet50.py --impl ours
run resnet_ours_cudnn 150 $TR 8 --master-addr 127.0.0.1 --master-port 29623 benchmarks/bench_resnet50.py --impl ours --no_tc
run resnet_torch 150 $TR 8 --master-addr 127.0.0.1 --master-port 29622 benchmarks/bench_resnet50.py --impl torch
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo done
