#!/bin/bash
# parity suite, cfg3 span sweep, then cfg3 / cfg2 / cfg5 lines.   usage: tools/c3_round.sh <tag>
TAG=${1:-rX}; O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest.log
bash tools/c3_sweep.sh
timeout 600 python bench.py --workload cfg3 --steps 100 --warmup 5 > $O/${TAG}_bench_cfg3.json 2> $O/${TAG}_bench_cfg3.err; echo "cfg3 rc=$?"; cut -c1-400 $O/${TAG}_bench_cfg3.json; tail -3 $O/${TAG}_bench_cfg3.err
timeout 600 python bench.py --no-cpu-baseline > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench.err; echo "cfg2 rc=$?"; cut -c1-330 $O/${TAG}_bench_cfg2.json
timeout 600 python bench.py --workload cfg5 --steps 50 --warmup 5 > $O/${TAG}_bench_cfg5_n1.json 2> $O/${TAG}_bench_cfg5.err; echo "cfg5 rc=$?"; cut -c1-330 $O/${TAG}_bench_cfg5_n1.json
