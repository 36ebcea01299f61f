#!/bin/bash
# quick A/B of a kernel change on one GPU: parity of the draw loop, then cfg2 / cfg3 / cfg5 / cfg1 step times
tag=${1:-quick}
python -m pytest -q -x -m gpu tests/test_gpu_parity.py tests/test_gpu_search2.py tests/test_gpu_full_size.py 2>&1 | tail -2
for wl in cfg2 cfg3 cfg5 cfg1; do
  python bench.py --workload $wl --steps 100 --warmup 5 --no-cpu-baseline --no-sharded > gpurun_out/${tag}_bench_$wl.json 2> gpurun_out/${tag}_bench_$wl.err
  python - $wl gpurun_out/${tag}_bench_$wl.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
print("%s: ms/step %.4f value %.4e e2e %.4e" % (sys.argv[1], d["ms_per_step"], d["value"], (d.get("e2e") or {}).get("value", 0)),
      "| search %.1f us draw %.1f us" % (d["roofline"].get("launch_ms", 0) * 1e3, (d["roofline"].get("integrate") or {}).get("launch_ms", 0) * 1e3) if "roofline" in d and d["roofline"].get("integrate") else "")
PY
done
