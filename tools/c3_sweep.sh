#!/bin/bash
# cfg3 rings-kernel launch-shape sweep (span = rings per work unit; slot-table bits): ms per step, flushed and L2-warm
O=gpurun_out; mkdir -p $O
for SPAN in ${SPANS:-1 2 3 4 6}; do
  for BITS in 13; do
    CS_TUNE_RING_SPAN=$SPAN CS_TUNE_RING_SLOT_BITS=$BITS timeout 300 python bench.py --workload cfg3 --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null \
      | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('span $SPAN bits $BITS: %.1f us flushed, %.1f us warm, launch %.1f us, checksum %d' % (d['ms_per_step']*1e3, d['replay_l2_warm']['ms_per_step']*1e3, d['roofline']['launch_ms']*1e3, d['map_checksum']))"
  done
done | tee $O/c3_sweep.txt
