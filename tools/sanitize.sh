#!/bin/bash
# compute-sanitizer over smoke() and a small parity subset.  Logs go to gpurun_out/<tag>_sanitizer_<tool>.log
# usage: tools/sanitize.sh <tag> [tools...]
# racecheck replays every shared-memory access and is ~100x slower than the others: it gets smoke() and one integration test.
tag=${1:-r2}
shift
tools=${@:-memcheck racecheck synccheck}
mkdir -p gpurun_out
subset="tests/test_gpu_parity.py::test_integrate_map_bit_exact tests/test_gpu_parity.py::test_search_edge_cases tests/test_gpu_parity.py::test_update_replay_bit_exact tests/test_gpu_search2.py::test_slab_batch_update_matches_oracle_per_session"
for t in $tools; do
  log=gpurun_out/${tag}_sanitizer_${t}.log
  tests=$subset
  [ $t = racecheck ] && tests="tests/test_gpu_parity.py::test_integrate_map_bit_exact"
  echo "== compute-sanitizer --tool $t: smoke()" > $log
  timeout 600 compute-sanitizer --tool $t --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" >> $log 2>&1
  echo "exit $?" >> $log
  echo "== compute-sanitizer --tool $t: $tests" >> $log
  timeout 600 compute-sanitizer --tool $t --print-limit 20 python -m pytest -q -x -m gpu $tests -p no:cacheprovider >> $log 2>&1
  echo "exit $?" >> $log
  tail -n 12 $log
done
