#!/bin/bash
# full ncu capture of one launch of the kernels matching $2 (regex), tag $1; further args are env settings
TAG=${1:-s2f}; K=${2:-cs_search2}; shift; shift
O=gpurun_out; mkdir -p $O
env "$@" timeout 900 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:$K -s 30 -c 1 -f -o $O/${TAG} python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
ls -la $O/${TAG}.ncu-rep
