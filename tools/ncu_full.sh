#!/bin/bash
# full ncu capture of one launch: tools/ncu_full.sh <tag> <kernel regex> <skip> <bench args...>
TAG=$1; K=$2; SKIP=$3; shift; shift; shift
O=gpurun_out; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o $O/${TAG} python bench.py "$@" > $O/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
