#!/bin/bash
# A/B of kernel variants built with tools/build_variant.sh: tools/variants_perf.sh [-w "cfg2 cfg3"] "<name>[:ENV=V,...]" ...
wls="cfg2 cfg3 cfg5"
if [ "$1" = "-w" ]; then wls=$2; shift; shift; fi
for spec in "$@"; do
  name=${spec%%:*}; envs=""
  [ "$spec" != "$name" ] && envs=$(echo ${spec#*:} | tr ',' ' ')
  for wl in $wls; do
    env CS_B200_LIB=$PWD/slam.net_b200/_build/variants/lib_$name.so $envs python bench.py --workload $wl --steps 100 --warmup 5 --no-cpu-baseline --no-sharded 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-28s %s: ms/step %.4f  e2e ms %.4f' % ('$spec', '$wl', d['ms_per_step'], (d.get('e2e') or {}).get('ms_per_step', 0)))"
  done
done
