#!/bin/bash
O=gpurun_out; mkdir -p $O; TAG=${1:-c5}; shift
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for cfg in "$@"; do
  n=$(echo $cfg | tr ' =' '__')
  env $cfg timeout 600 python bench.py --workload cfg5 --steps 10 --warmup 2 > $O/${TAG}_cfg5_$n.json 2> $O/${TAG}_cfg5_$n.err
  echo "$cfg rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_cfg5_$n.json").read().strip().splitlines()[-1])
    print("  ms/step %.4f value %.3e sessions/s %.1f"%(d["ms_per_step"],d["value"],d.get("sessions_per_s",0)))
except Exception as e:
    print("  parse failed", e); print(open("$O/${TAG}_cfg5_$n.err").read()[-1500:])
PY
done
