#!/bin/bash
# round-2 GPU check K: ring first (r2i saw 8.5 ms/step after an ncu run: host state or code?), tests of the RING changes, tiled, cfg2
O=gpurun_out/r2k; mkdir -p $O
timeout 300 python bench.py --workload ring --steps 10 --warmup 3 > $O/bench_ring.json 2> $O/bench_ring.err; echo "ring rc=$?"; head -c 300 $O/bench_ring.json; echo
timeout 900 python -m pytest tests/test_radon_gpu.py tests/test_gicp_gpu.py tests/test_elevation_gpu.py -q 2>&1 | tail -3
timeout 300 python scripts/prof_gicp.py 8 2 2>&1 | grep -E "per-kernel" | tail -1
timeout 400 python bench.py --workload tiled --steps 5 --warmup 3 > $O/bench_tiled.json 2> $O/bench_tiled.err; echo "tiled rc=$?"; head -c 300 $O/bench_tiled.json; echo
timeout 400 python bench.py --steps 10 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "cfg2 rc=$?"; head -c 300 $O/bench_cfg2.json; echo; tail -3 $O/*.err
