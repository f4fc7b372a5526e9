#!/bin/bash
# One multi-GPU session on the box: correctness first, then every multi-GPU number we report.
#   gpurun --gpus N -- bash benchmarks/run_multi.sh N
N=${1:-8}
OUT=gpurun_out/n$N
mkdir -p $OUT
export NCCL_DEBUG=WARN
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2> $OUT/$name.err; echo "rc=$? $(tail -c 400 $OUT/$name.log | tr '\n' ' ' | cut -c1-400)"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node"
run tests 420 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 200 -s
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  if [ $n -eq 1 ]; then
    run bench_ours_n1 120 python bench.py --steps 200 --warmup 10
    run bench_ref_n1 200 python bench.py --impl reference --steps 60 --warmup 5
  else
    run bench_ours_n$n 150 $TR $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 200 --warmup 10
    run bench_ref_n$n 240 $TR $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --impl reference --steps 60 --warmup 5
  fi
done
run sweep_n$N 240 $TR $N --master-addr 127.0.0.1 --master-port 29620 benchmarks/allreduce_sweep.py --out $OUT/allreduce_sweep.json
run resnet_ours 240 $TR $N --master-addr 127.0.0.1 --master-port 29621 benchmarks/bench_resnet50.py --impl ours
run resnet_torch 240 $TR $N --master-addr 127.0.0.1 --master-port 29622 benchmarks/bench_resnet50.py --impl torch
NCCL_DEBUG=INFO timeout 60 $TR 2 --master-addr 127.0.0.1 --master-port 29623 benchmarks/allreduce_sweep.py --iters 1 2>&1 | grep -iE "nvls|multicast" | head -5 > $OUT/nccl_nvls.txt
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo done
