#!/bin/bash
# round-2 GPU check A: parity tests, the three bench workloads and their reference arms on one B200
O=gpurun_out/r2a; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "cfg2 rc=$?"
timeout 300 python bench.py --workload ring --steps 10 --warmup 3 > $O/bench_ring.json 2> $O/bench_ring.err; echo "ring rc=$?"
timeout 400 python bench.py --workload tiled --steps 5 --warmup 3 > $O/bench_tiled.json 2> $O/bench_tiled.err; echo "tiled rc=$?"
( time timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/ref_cfg2.json 2> $O/ref_cfg2.err ) 2> $O/ref_cfg2.time; echo "ref cfg2 rc=$?"
( time timeout 300 python bench.py --impl reference --workload ring --steps 2 --warmup 1 > $O/ref_ring.json 2> $O/ref_ring.err ) 2> $O/ref_ring.time
( time timeout 300 python bench.py --impl reference --workload tiled --steps 2 --warmup 1 > $O/ref_tiled.json 2> $O/ref_tiled.err ) 2> $O/ref_tiled.time
tail -3 $O/pytest.log; head -c 600 $O/bench_cfg2.json; echo; head -c 400 $O/bench_ring.json; echo; head -c 400 $O/bench_tiled.json; echo; tail -5 $O/*.err
