#!/bin/bash
# cfg4 (8192^2 map, 65536 candidates) on one GPU: warp-per-candidate vs slab search
O=gpurun_out; mkdir -p $O; TAG=${1:-c4}
for cfg in "CS_TUNE_SEARCH2=-1" "CS_NONE=1"; do
  n=$(echo $cfg | tr ' =' '__')
  env $cfg timeout 600 python bench.py --workload cfg4 --steps 100 --warmup 10 > $O/${TAG}_cfg4_$n.json 2> $O/${TAG}_cfg4_$n.err
  echo "$cfg rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_cfg4_$n.json").read().strip().splitlines()[-1])
    print("  ms/step %.4f value %.3e e2e %.3e launches %s pose %s"%(d["ms_per_step"],d["value"],d["e2e"]["value"],d["gpu_launches"],d["final_pose"]))
except Exception as e:
    print("  parse failed", e); print(open("$O/${TAG}_cfg4_$n.err").read()[-1500:])
PY
done
