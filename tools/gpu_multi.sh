#!/bin/bash
# 2-GPU round: multi-GPU tests, then cfg2 (one replay per GPU), cfg4 (candidate split) and cfg5 (session batches) at N=2
TAG=${1:-rXm}; N=${2:-2}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > $O/${TAG}_pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest_multi.log
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N"
timeout 600 $R --no-cpu-baseline > $O/${TAG}_bench_cfg2_n$N.json 2> $O/${TAG}_cfg2.err; echo "cfg2 rc=$?"; tail -1 $O/${TAG}_bench_cfg2_n$N.json | cut -c1-400
timeout 600 $R --workload cfg4 --steps 100 --warmup 10 > $O/${TAG}_bench_cfg4_n$N.json 2> $O/${TAG}_cfg4.err; echo "cfg4 rc=$?"; tail -1 $O/${TAG}_bench_cfg4_n$N.json | cut -c1-400
timeout 900 $R --workload cfg5 --steps 20 --warmup 3 > $O/${TAG}_bench_cfg5_n$N.json 2> $O/${TAG}_cfg5.err; echo "cfg5 rc=$?"; tail -1 $O/${TAG}_bench_cfg5_n$N.json | cut -c1-400
timeout 900 python bench.py --workload cfg5 --steps 20 --warmup 3 > $O/${TAG}_bench_cfg5_n1.json 2> $O/${TAG}_cfg5_1.err; echo "cfg5 n1 rc=$?"; tail -1 $O/${TAG}_bench_cfg5_n1.json | cut -c1-300
