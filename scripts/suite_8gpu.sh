#!/usr/bin/env bash
# Round-end evidence on one 8 x B200 node (gpurun --gpus 8 -- bash scripts/suite_8gpu.sh r03): the 8-rank NCCL
# parity test, config 3 (global batch 32768) and config 4 (D=768, DINOv2-L dim 1024, global batch 65536) bench lines
# at 8 ranks and the config-5 harness with the drop-in loss.  Every step has its own short timeout (an 8-GPU minute
# costs eight).  Outputs: gpurun_out/<tag>_*.  N = 2 / 4: `gpurun --gpus N -- bash scripts/bench_scaling.sh`.
set -x
tag=${1:-rXX}
mkdir -p gpurun_out
run() { n=$1; shift; timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
timeout 200 python -m pytest tests/test_gpu_dist.py -m gpu -q -k eight_gpu 2>&1 | tail -4 > gpurun_out/${tag}_dist8_tests.log; cat gpurun_out/${tag}_dist8_tests.log
run 8 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n8.json 2> gpurun_out/${tag}_n8.err; tail -c 200 gpurun_out/${tag}_bench_n8.json
run 8 bench.py --gpus 8 --steps 10 --warmup 3 --clip-dim 768 --dino-dim 1024 --no-head --batch 65536 > gpurun_out/${tag}_bench_c4_n8_b65536.json 2> gpurun_out/${tag}_c4n8.err
run 8 scripts/train_step_harness.py --loss ours --batch 512 --steps 6 > gpurun_out/${tag}_c5_ours_n8.json 2> gpurun_out/${tag}_c5o8.err
cut -c1-300 gpurun_out/${tag}_c5_ours_n8.json
for f in bench_n8 bench_c4_n8_b65536; do python -c "import json; d=json.load(open('gpurun_out/${tag}_$f.json')); print('$f', d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])"; done
