timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_robustness.py tests/test_gpu_selftest.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_ll.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/plain.log").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],3), d["step_roofline"]["serial_ms_per_step"], d["clocks"]["sm_mhz"], d["loss"])
PY
