#!/usr/bin/env bash
# One bench line at N GPUs of one node, as the driver launches it: `gpurun --gpus N -- bash scripts/bench_scaling.sh N tag`
set -x
n=${1:-2}; tag=${2:-rXX}
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $n > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/bench_n$n.err
tail -c 300 gpurun_out/${tag}_bench_n$n.json
