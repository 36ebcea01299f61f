#!/bin/bash
# parity suite + smoke, cfg3 span check, cfg3 line + ncu captures of the several-rounds rings kernel, cfg5 at 128 sessions
TAG=${1:-rX}; O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/${TAG}_smoke.log
SPANS="1 2 3" bash tools/c3_sweep.sh
timeout 600 python bench.py --workload cfg3 --steps 100 --warmup 5 > $O/${TAG}_bench_cfg3.json 2> $O/${TAG}_bench_cfg3.err; echo "cfg3 rc=$?"; cut -c1-330 $O/${TAG}_bench_cfg3.json; tail -3 $O/${TAG}_bench_cfg3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${TAG}_launches_cfg3.csv python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_launch_cfg3.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cs_rings_kernel -s 6 -c 1 -f -o $O/${TAG}_rings_cfg3 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_rings_cfg3.log 2>&1; echo "ncu rings rc=$?"
timeout 300 python bench.py --workload cfg5 --sessions 128 --steps 20 --warmup 3 > $O/${TAG}_bench_cfg5_128.json 2>/dev/null; cut -c1-300 $O/${TAG}_bench_cfg5_128.json
