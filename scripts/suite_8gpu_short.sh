#!/usr/bin/env bash
# Short 8 x B200 check (gpurun --gpus 8 -- bash scripts/suite_8gpu_short.sh r04): the 8-rank NCCL parity test, the
# config-3 bench line (global batch 32768) and the config-4 line (D=768, DINOv2-L dim 1024, global batch 65536).
set -x
tag=${1:-rXX}
mkdir -p gpurun_out
run() { n=$1; shift; timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
timeout 200 python -m pytest tests/test_gpu_dist.py -m gpu -q -k eight_gpu 2>&1 | tail -4 > gpurun_out/${tag}_dist8_tests.log; cat gpurun_out/${tag}_dist8_tests.log
run 8 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_n8.json 2> gpurun_out/${tag}_n8.err
run 8 bench.py --gpus 8 --steps 10 --warmup 3 --clip-dim 768 --dino-dim 1024 --no-head --batch 65536 --no-cpu-baseline > gpurun_out/${tag}_bench_c4_n8_b65536.json 2> gpurun_out/${tag}_c4n8.err
for f in bench_n8 bench_c4_n8_b65536; do python -c "
import json; d=json.loads(open('gpurun_out/${tag}_$f.json').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['value'], d['e2e']['value'], d['step_roofline']['serial_ms_per_step'], d['step_roofline']['tile_kernel_ms_per_serial_step'], {k: v['ms'] for k, v in d['kernels'].items()}, d['clocks'])"; done
tail -n 3 gpurun_out/${tag}_n8.err gpurun_out/${tag}_c4n8.err | cut -c1-400
