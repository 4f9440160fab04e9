R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 scripts/comm_bench.py"
timeout 200 $R 2>&1 | grep "world="
NCCL_PROTO=Simple timeout 200 $R 2>&1 | grep "world=" | sed 's/^/PROTO=Simple /'
NCCL_ALGO=NVLS timeout 200 $R 2>&1 | grep "world=" | sed 's/^/ALGO=NVLS /'
nproc
