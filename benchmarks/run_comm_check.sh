#!/bin/bash
# Communication-layer session: gpurun --gpus N -- bash benchmarks/run_comm_check.sh N
N=${1:-2}
OUT=gpurun_out/comm_n$N
mkdir -p $OUT
export NCCL_DEBUG=WARN
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2> $OUT/$name.err; echo "rc=$? $(tail -c 600 $OUT/$name.log | tr '\n' ' ' | cut -c1-600)"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node"
run tests_comm 300 python -m pytest tests/test_gpu_comm.py tests/test_gpu_toynet.py -m gpu -q -x --timeout 120
run tests_multi 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 240 -s
run sweep 300 $TR $N --master-addr 127.0.0.1 --master-port 29620 benchmarks/allreduce_sweep.py --out $OUT/allreduce_sweep.json
run bench_n1 150 python bench.py --steps 200 --warmup 10
run bench_f32 200 $TR $N --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus $N --steps 200 --warmup 10
run bench_bf16 200 $TR $N --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus $N --steps 200 --warmup 10 --comm_wire bf16 --census $OUT/census_n${N}_bf16.json
run bench_f32_census 200 $TR $N --master-addr 127.0.0.1 --master-port 29603 bench.py --gpus $N --steps 100 --warmup 10 --census $OUT/census_n${N}_f32.json
run bench_pg 200 $TR $N --master-addr 127.0.0.1 --master-port 29604 bench.py --gpus $N --steps 100 --warmup 10 --comm pg
echo done
