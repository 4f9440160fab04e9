#!/usr/bin/env bash
# Round-end evidence on one 8 x B200 node (gpurun --gpus 8 -- bash scripts/suite_8gpu.sh r03): NCCL parity tests,
# config 3 (global batch 32768) and config 4 (D=768, DINOv2-L dim 1024, global batch 65536) bench lines at 8 ranks,
# config 3 at 2 and 4 ranks, and the config-5 harness with both losses.  Outputs: gpurun_out/<tag>_*.
set -x
tag=${1:-rXX}
mkdir -p gpurun_out
run() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -4 > gpurun_out/${tag}_dist_tests.log; cat gpurun_out/${tag}_dist_tests.log
run 8 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n8.json 2> gpurun_out/${tag}_n8.err; tail -c 200 gpurun_out/${tag}_bench_n8.json
run 8 bench.py --gpus 8 --steps 10 --warmup 3 --clip-dim 768 --dino-dim 1024 --no-head --batch 65536 > gpurun_out/${tag}_bench_c4_n8_b65536.json 2> gpurun_out/${tag}_c4n8.err
run 8 scripts/train_step_harness.py --loss reference --batch 512 --steps 6 > gpurun_out/${tag}_c5_reference_n8.json 2> gpurun_out/${tag}_c5r8.err
run 8 scripts/train_step_harness.py --loss ours --batch 512 --steps 6 > gpurun_out/${tag}_c5_ours_n8.json 2> gpurun_out/${tag}_c5o8.err
cat gpurun_out/${tag}_c5_reference_n8.json gpurun_out/${tag}_c5_ours_n8.json | cut -c1-300
run 4 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n4.json 2> gpurun_out/${tag}_n4.err
run 2 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_n2.err
for f in n2 n4 n8; do python -c "import json; d=json.load(open('gpurun_out/${tag}_bench_$f.json')); print('$f', d['ms_per_step'], d['value'], d['e2e']['value'])"; done
