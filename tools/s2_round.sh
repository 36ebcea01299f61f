#!/bin/bash
# slab-search bring-up: its parity tests, then A/B bench runs and a launch-shape sweep (outputs under gpurun_out/<tag>_*)
# usage: tools/s2_round.sh <tag> "<env settings>" "<env settings>" ...   ("-" = no setting)
TAG=${1:-s2}; shift
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_search2.py -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/${TAG}_pytest.log
B="python bench.py --steps 300 --warmup 20 --no-cpu-baseline"
for cfg in "$@"; do
  n=$(echo $cfg | tr ' =' '__')
  if [ "$cfg" = "-" ]; then cfg="CS_NONE=1"; fi
  env $cfg timeout 300 $B > $O/${TAG}_bench_$n.json 2> $O/${TAG}_bench_$n.err
  echo "$cfg rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_bench_$n.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("  ms/step %.4f value %.3e e2e %.3e p50 %.4f search_ms %.4f rings_ms %.4f launches %s"%(d["ms_per_step"],d["value"],d["e2e"]["value"],d["e2e"]["scan_to_pose_latency_ms"]["p50"],r["launch_ms"],r["integrate"]["launch_ms"],d["gpu_launches"]))
except Exception as e:
    print("  parse failed", e); print(open("$O/${TAG}_bench_$n.err").read()[-1500:])
PY
done
