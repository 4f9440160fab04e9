timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 > gpurun_out/r01c_bench_n4.json 2> gpurun_out/bench_n4.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r01c_bench_n4.json").read().strip().splitlines()[-1])
print("N=4", round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], round(d["value"]), round(d["e2e"]["value"]))
PY
