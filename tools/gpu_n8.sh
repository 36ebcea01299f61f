#!/bin/bash
# 8-GPU lines: cfg5 (session batches, the scaling config) and cfg2 (one replay per GPU)
TAG=${1:-rXm}; N=${2:-8}
O=gpurun_out; mkdir -p $O
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N"
timeout 600 $R --workload cfg5 --steps 20 --warmup 3 > $O/${TAG}_bench_cfg5_n$N.json 2> $O/${TAG}_cfg5.err; echo "cfg5 rc=$?"; tail -1 $O/${TAG}_bench_cfg5_n$N.json | cut -c1-400
timeout 600 $R --no-cpu-baseline > $O/${TAG}_bench_cfg2_n$N.json 2> $O/${TAG}_cfg2.err; echo "cfg2 rc=$?"; tail -1 $O/${TAG}_bench_cfg2_n$N.json | cut -c1-400
