#!/bin/bash
# launch-shape sweeps (diagnostics): slab search (points per cluster x candidates per slab) and wedge kernel (blocks, ring split)
out=${1:-gpurun_out/r2_sweep.txt}
: > $out
run() { python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-sharded 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-40s step %.2f us  search %.2f  integrate %.2f  e2e %.3e  warm %.2f' % ('$1', d['ms_per_step']*1e3, d['roofline']['launch_ms']*1e3, d['roofline']['integrate']['launch_ms']*1e3, d['e2e']['value'], d['replay_l2_warm']['ms_per_step']*1e3))" >> $out; }
run default
for p in 32 64 128; do for t in 128 256 480; do CS_TUNE_S2_POINTS=$p CS_TUNE_S2_THREADS=$t run "s2 points=$p threads=$t"; done; done
for b in 296 444 592; do for s in 2 4 8; do CS_TUNE_W_BLOCKS=$b CS_TUNE_W_SUB=$s run "wedge blocks=$b sub=$s"; done; done
cat $out
