#!/bin/bash
# round-2 GPU check B: the v2 kNN kernel and the one-pass elevation replay -- parity, statistics, step time, ncu
O=gpurun_out/r2b; mkdir -p $O
timeout 900 python -m pytest tests/test_gicp_gpu.py tests/test_elevation_gpu.py tests/test_cabi.py -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 $O/pytest.log
timeout 300 python scripts/prof_gicp.py 8 2 > $O/prof_v2.log 2>&1; echo "prof rc=$?"; cat $O/prof_v2.log
MRSLAM_KNN_V1=1 timeout 300 python scripts/prof_gicp.py 8 1 > $O/prof_v1.log 2>&1; grep "per-kernel" $O/prof_v1.log
timeout 400 python bench.py --steps 10 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "cfg2 rc=$?"; head -c 300 $O/bench_cfg2.json; echo; tail -3 $O/bench_cfg2.err
timeout 400 python bench.py --workload tiled --steps 5 --warmup 3 > $O/bench_tiled.json 2> $O/bench_tiled.err; echo "tiled rc=$?"; head -c 300 $O/bench_tiled.json; echo; tail -3 $O/bench_tiled.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knn_search_v2 -s 2 -c 1 -o $O/r2_knn_search_v2 python scripts/prof_gicp.py 8 1 > $O/ncu_knn.log 2>&1; echo "ncu rc=$?"; tail -3 $O/ncu_knn.log
