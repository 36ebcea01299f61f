#!/bin/bash
# 8-GPU cfg5 line only (session batches, strong scaling over the fixed 1024 sessions)
TAG=${1:-rXm}; N=${2:-8}
O=gpurun_out; mkdir -p $O
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N"
timeout 300 $R --workload cfg5 --steps 20 --warmup 3 > $O/${TAG}_bench_cfg5_n$N.json 2> $O/${TAG}_cfg5.err; echo "cfg5 rc=$?"; tail -1 $O/${TAG}_bench_cfg5_n$N.json | cut -c1-300
