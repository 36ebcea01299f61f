#!/bin/bash
# N-GPU lines for cfg2 (one replay per GPU), cfg5 (session batches) and cfg4 (candidate split)
TAG=${1:-rXs}; N=${2:-8}
O=gpurun_out; mkdir -p $O
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N"
timeout 600 $R --no-cpu-baseline --steps 300 > $O/${TAG}_bench_cfg2_n$N.json 2> $O/${TAG}_cfg2_n$N.err; echo "cfg2 rc=$?"; tail -1 $O/${TAG}_bench_cfg2_n$N.json | cut -c1-200
timeout 900 $R --workload cfg5 --steps 20 --warmup 3 > $O/${TAG}_bench_cfg5_n$N.json 2> $O/${TAG}_cfg5_n$N.err; echo "cfg5 rc=$?"; tail -1 $O/${TAG}_bench_cfg5_n$N.json | cut -c1-200
timeout 600 $R --workload cfg4 --steps 100 --warmup 10 > $O/${TAG}_bench_cfg4_n$N.json 2> $O/${TAG}_cfg4_n$N.err; echo "cfg4 rc=$?"; tail -1 $O/${TAG}_bench_cfg4_n$N.json | cut -c1-200
