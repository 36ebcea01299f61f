#!/bin/bash
# last call of a session: parity suite, smoke, both arms of the default line
TAG=${1:-rX}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/${TAG}_smoke.log
timeout 600 python bench.py --impl reference --steps 40 --warmup 5 > $O/${TAG}_bench_cfg2_reference.json 2>$O/${TAG}_ref.err; echo "ref rc=$?"; cut -c1-200 $O/${TAG}_bench_cfg2_reference.json
timeout 600 python bench.py > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cat $O/${TAG}_bench_cfg2.json; tail -3 $O/${TAG}_bench.err
