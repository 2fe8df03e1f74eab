#!/bin/bash
# round-2 GPU check F: kNN v2 (one warp per tile, two-phase scan): parity, chunk/cap variants, ncu
O=gpurun_out/r2f; mkdir -p $O
export MRSLAM_KNN_V2=1
timeout 600 python -m pytest tests/test_gicp_gpu.py -q -x > $O/pytest_gicp.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gicp.log
export MRSLAM_KNN_V2=1
for v in default ch16 ch32 ch24cap64 ch32cap64; do
  if [ $v = default ]; then unset MRSLAM_LIB; else export MRSLAM_LIB=$PWD/build/variants/libmrslam_$v.so; fi
  echo "== variant $v"; timeout 300 python scripts/prof_gicp.py 8 2 2>&1 | grep -E "knn v2|per-kernel" | tail -3
done
unset MRSLAM_LIB
env -u MRSLAM_KNN_V2 MRSLAM_KNN_V1=1 timeout 300 python scripts/prof_gicp.py 8 1 2>&1 | grep "per-kernel" | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_knn -c 8 --csv --log-file $O/knn_launches.csv python scripts/prof_gicp.py 8 1 > $O/l.log 2>&1; grep -v "^==" $O/knn_launches.csv | awk -F'","' '{print $5, $13, $15}' | head -20
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_knn_search_v2 -s 2 -c 1 -o $O/r2_knn_search_v2b python scripts/prof_gicp.py 8 1 > $O/ncu_knn.log 2>&1; echo "ncu rc=$?"
timeout 400 python bench.py --steps 10 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "cfg2 rc=$?"; head -c 250 $O/bench_cfg2.json; echo; tail -3 $O/bench_cfg2.err
# warps-per-block A/B of the two round-1 search kernels (v1 kNN, NN-1)
for cfg in "4 4" "1 4" "4 1" "1 1"; do set -- $cfg
  echo "== KNN_WPB=$1 NN1_WPB=$2"; env -u MRSLAM_KNN_V2 MRSLAM_KNN_WPB=$1 MRSLAM_NN1_WPB=$2 timeout 300 python scripts/prof_gicp.py 8 2 2>&1 | grep -E "per-kernel" | tail -1
done
env -u MRSLAM_KNN_V2 MRSLAM_KNN_WPB=1 MRSLAM_NN1_WPB=1 timeout 600 python -m pytest tests/test_gicp_gpu.py -q -x 2>&1 | tail -2
# own radix sort + one-pass replay
env -u MRSLAM_KNN_V2 timeout 900 python -m pytest tests/test_elevation_gpu.py -q 2>&1 | tail -3
env -u MRSLAM_KNN_V2 timeout 400 python bench.py --workload tiled --steps 5 --warmup 3 > $O/bench_tiled.json 2> $O/bench_tiled.err; echo "tiled rc=$?"; head -c 250 $O/bench_tiled.json; echo; tail -3 $O/bench_tiled.err
env -u MRSLAM_KNN_V2 timeout 400 python bench.py --steps 10 --warmup 3 > $O/bench_cfg2_v1.json 2> $O/bench_cfg2_v1.err; echo "cfg2(v1 knn) rc=$?"; head -c 250 $O/bench_cfg2_v1.json; echo; tail -3 $O/bench_cfg2_v1.err
