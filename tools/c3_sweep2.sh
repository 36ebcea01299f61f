#!/bin/bash
# cfg3 rings-kernel launch-shape sweep: threads per block x slot-table bits x blocks per SM x rings per unit
O=gpurun_out; mkdir -p $O
run() {
  CS_TUNE_RING_THREADS=$1 CS_TUNE_RING_SLOT_BITS=$2 CS_TUNE_RING_BLOCKS_PER_SM=$3 CS_TUNE_RING_SPAN=$4 timeout 300 python bench.py --workload cfg3 --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('threads $1 bits $2 blocks/SM $3 span $4: %.1f us flushed, %.1f us warm, launch %.1f us, checksum %d' % (d['ms_per_step']*1e3, d['replay_l2_warm']['ms_per_step']*1e3, d['roofline']['launch_ms']*1e3, d['map_checksum']))"
}
{
run 512 13 2 2
run 512 12 2 2
run 256 13 3 2
run 256 12 4 2
run 256 12 4 1
run 256 12 4 4
run 256 11 4 2
run 384 12 2 2
run 128 12 4 2
run 128 11 8 2
} | tee $O/c3_sweep2.txt
