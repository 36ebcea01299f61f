#!/bin/bash
# usage: tools/sweep.sh <tag> "ENV1=a ENV2=b" "ENV1=c" ...   -- one short bench per environment setting
TAG=$1; shift
O=gpurun_out; mkdir -p $O
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'):
        print(line); continue
    d=json.loads(line)
    r=d['roofline']
    print('step %.2f us  warm %.2f us  search %.2f us  rings %.2f us  e2e %.2f us  p50 %.1f us  gather frac %.3f  value %.3e' % (
        d['ms_per_step']*1e3, d['replay_l2_warm']['ms_per_step']*1e3, r['launch_ms']*1e3, r['integrate']['launch_ms']*1e3,
        d['e2e']['ms_per_step']*1e3, d['e2e']['scan_to_pose_latency_ms']['p50']*1e3, r['gather']['frac'] or 0, d['value']))
" | tee -a $O/${TAG}_sweep.txt
done
