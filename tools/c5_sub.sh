#!/bin/bash
# cfg5 at the per-GPU load of an 8-GPU run (128 sessions) and of a 1-GPU run (1024): one batch vs several on their own streams
O=gpurun_out; mkdir -p $O
run() {
  timeout 300 python bench.py --workload cfg5 --sessions $1 --sub-batches $2 --steps 20 --warmup 3 2>$O/c5_sub.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sessions $1 sub-batches $2: %.3f ms per step, %.0f sessions/s, launches %d' % (d['ms_per_step'], d['sessions_per_s'], d['gpu_launches']))" || tail -3 $O/c5_sub.err
}
{
run 128 1
run 128 2
run 128 4
run 256 1
run 256 2
run 512 1
run 512 2
run 1024 1
run 1024 2
} | tee $O/c5_sub.txt
