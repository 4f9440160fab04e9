timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_selftest.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/tri.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/tri.log").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],3), d["step_roofline"]["serial_ms_per_step"], d["clocks"]["sm_mhz"], d["loss"], round(d["value"]))
for k,v in d["kernels"].items(): print("   ",k,v["ms"],v["exec_tflops"],v["launches"])
PY
