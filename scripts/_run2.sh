timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for c in 1 0; do
DSOFT_CONCURRENCY=$c timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1 conc=$c', round(d['ms_per_step'],4), d['step_roofline']['serial_ms_per_step'], d['clocks']['sm_mhz'], round(d['value']))"
done
for c in 1 0; do
DSOFT_CONCURRENCY=$c timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=2 conc=$c', round(d['ms_per_step'],4), d['step_roofline']['serial_ms_per_step'], d['clocks']['sm_mhz'], round(d['value']))"
done
