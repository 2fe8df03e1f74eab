#!/bin/bash
# round-2 GPU check O: the whole parity suite + smoke() + the driver's default commands on the final tree
O=gpurun_out/r2o; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench_default.time; echo "bench default rc=$?"; head -c 300 $O/bench_default.json; echo; grep real $O/bench_default.time
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/ref_default.json 2> $O/ref_default.err ) 2> $O/ref_default.time; echo "ref rc=$?"; head -c 300 $O/ref_default.json; echo; grep real $O/ref_default.time
