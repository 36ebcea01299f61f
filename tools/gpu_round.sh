#!/bin/bash
# One GPU-box call: parity suite, smoke, bench (both arms, streaming replay, cfg4), ncu launch list and full captures of the kernels.
# usage: tools/gpu_round.sh <tag>   (outputs under gpurun_out/<tag>_*)
TAG=${1:-rX}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/${TAG}_smoke.log
timeout 600 python bench.py > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cat $O/${TAG}_bench_cfg2.json; tail -3 $O/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 40 --warmup 5 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "ref rc=$?"; cat $O/${TAG}_bench_ref.json
timeout 600 python bench.py --streaming --no-cpu-baseline > $O/${TAG}_bench_cfg2_streaming.json 2> $O/${TAG}_bench_streaming.err; echo "streaming rc=$?"; cat $O/${TAG}_bench_cfg2_streaming.json
CS_TUNE_SEARCH2=-1 timeout 600 python bench.py --no-cpu-baseline > $O/${TAG}_bench_cfg2_warpsearch.json 2> $O/${TAG}_bench_warp.err; echo "warp-search rc=$?"
timeout 600 python bench.py --workload cfg4 --steps 100 --warmup 10 > $O/${TAG}_bench_cfg4_n1.json 2> $O/${TAG}_bench_cfg4.err; echo "cfg4 rc=$?"; cat $O/${TAG}_bench_cfg4_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_cfg2.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
for K in cs_search2 cs_sort cs_rings; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${K}_kernel -s 30 -c 2 -f -o $O/${TAG}_${K#cs_} python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_${K#cs_}.log 2>&1; echo "ncu $K rc=$?"
done
ls -la $O | tail -30
