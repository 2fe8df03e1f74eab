#!/bin/bash
# round-2 GPU check G: warps-per-block A/B for the two search kernels (block slots freed per warp instead of per 4 warps)
O=gpurun_out/r2g; mkdir -p $O
for cfg in "4 4" "1 4" "4 1" "1 1"; do set -- $cfg
  echo "== KNN_WPB=$1 NN1_WPB=$2"; MRSLAM_KNN_WPB=$1 MRSLAM_NN1_WPB=$2 timeout 300 python scripts/prof_gicp.py 8 2 2>&1 | grep -E "per-kernel" | tail -1
done
timeout 600 python -m pytest tests/test_gicp_gpu.py -q -x 2>&1 | tail -2
MRSLAM_KNN_WPB=1 MRSLAM_NN1_WPB=1 timeout 600 python -m pytest tests/test_gicp_gpu.py -q -x 2>&1 | tail -2
