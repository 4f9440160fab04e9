timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > gpurun_out/r01c_bench_n8.json 2> gpurun_out/bench_n8.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r01c_bench_n8.json").read().strip().splitlines()[-1])
print("N=8", round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], round(d["value"]), round(d["e2e"]["value"]))
for k,v in d["kernels"].items(): print("   ",k,v["ms"],v["exec_tflops"],v["launches"])
PY
