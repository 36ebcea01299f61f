#!/bin/bash
# cfg5 at small per-GPU loads: rings per block (span) and slot-table bits of the batch rings kernel
O=gpurun_out; mkdir -p $O
run() {
  CS_TUNE_RING_SPAN=$2 CS_TUNE_RING_SLOT_BITS=$3 timeout 300 python bench.py --workload cfg5 --sessions $1 --steps 20 --warmup 3 2>$O/c5_span.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sessions $1 span $2 bits $3: %.3f ms per step, %.0f sessions/s' % (d['ms_per_step'], d['sessions_per_s']))" || tail -3 $O/c5_span.err
}
{
run 128 64 11
run 128 32 11
run 128 16 11
run 128 32 10
run 128 16 10
run 128 8 10
run 256 32 11
run 256 16 10
run 512 32 11
run 1024 32 11
} | tee $O/c5_span.txt
