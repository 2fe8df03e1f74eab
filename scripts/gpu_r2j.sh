#!/bin/bash
# round-2 GPU check J (2 GPUs): quick single-GPU parity/timing of the k_cov split and the NN-1 visit scheme, then the N=2 bench legs
O=gpurun_out/r2j; mkdir -p $O
timeout 600 python -m pytest tests/test_gicp_gpu.py -q 2>&1 | tail -2
echo "== nn1 wpq (default)"; timeout 300 python scripts/prof_gicp.py 8 2 2>&1 | grep -E "per-kernel|nn1 pair evals" | tail -2
echo "== nn1 lpq (variant)"; MRSLAM_LIB=$PWD/build/variants/libmrslam_nn1lpq.so timeout 300 python scripts/prof_gicp.py 8 2 2>&1 | grep -E "per-kernel|nn1 pair evals" | tail -2
bash scripts/gpu_r2d.sh 2
cp gpurun_out/r2d_n2/* $O/ 2>/dev/null
