#!/usr/bin/env bash
# Round-end evidence on ONE B200 (gpurun -- bash scripts/suite_1gpu.sh r03): the GPU test suite, the default bench
# line, the config-2 lines (eager and CUDA-graphed), the config-4 dims line, the config-5 harness with both losses,
# then the ncu launch list and the full-metric capture of one step's kernels.  Outputs: gpurun_out/<tag>_*.
set -x
tag=${1:-rXX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${tag}_gpu_tests.log; cat gpurun_out/${tag}_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; tail -c 300 gpurun_out/${tag}_bench_n1.json
timeout 300 python bench.py --batch 4096 --steps 50 --warmup 10 --no-cpu-baseline --no-graph > gpurun_out/${tag}_bench_c2_b4096.json 2> gpurun_out/${tag}_c2.err
timeout 300 python bench.py --batch 4096 --steps 50 --warmup 10 --no-cpu-baseline --graph > gpurun_out/${tag}_bench_c2_b4096_graphed.json 2> gpurun_out/${tag}_c2g.err
timeout 600 python bench.py --clip-dim 768 --dino-dim 1024 --no-head --batch 65536 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_bench_c4dims_1gpu.json 2> gpurun_out/${tag}_c4.err
timeout 300 python scripts/train_step_harness.py --loss reference --batch 512 --steps 6 > gpurun_out/${tag}_c5_reference_n1.json 2> gpurun_out/${tag}_c5r.err
timeout 300 python scripts/train_step_harness.py --loss ours --batch 512 --steps 6 > gpurun_out/${tag}_c5_ours_n1.json 2> gpurun_out/${tag}_c5o.err
cat gpurun_out/${tag}_c5_reference_n1.json gpurun_out/${tag}_c5_ours_n1.json | cut -c1-400
timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_ll.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dsoft_ -s 32 -c 8 -o gpurun_out/${tag}_tiles \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
