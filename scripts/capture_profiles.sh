#!/usr/bin/env bash
# Everything behind profiles/rNN_*: run on ONE B200 (e.g. `gpurun -- bash scripts/capture_profiles.sh r01c`).
# 1. smoke() and the default bench line (not under a profiler);
# 2. ncu launch list of a short bench run (cold-cache, serialised: shares of the step, not absolute times);
# 3. ncu --set full of one launch of every tile kernel of the third step (skip 36 = 4 steps x 9 launches).
# Outputs go to gpurun_out/; copy the summaries you want to keep into profiles/ (see DESIGN.md section 4).
set -x
tag=${1:-rXX}
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 400 gpurun_out/${tag}_bench_n1.json
timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_ll.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dsoft_ -s 36 -c 9 -o gpurun_out/${tag}_tiles \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
