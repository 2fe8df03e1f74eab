#!/bin/bash
# round-2 GPU check U: micro-variants of the search / covariance kernels (launch bounds, prefetch depth, super-tile group, home radius)
for v in default knn6 knn10 pf1 pf4 sg1 sg4 home0 cov4 nn1b10; do
  if [ $v = default ]; then unset MRSLAM_LIB; else export MRSLAM_LIB=$PWD/build/variants/libmrslam_$v.so; fi
  echo "== $v"; timeout 300 python scripts/prof_gicp.py 8 2 2>&1 | grep -E "per-kernel" | tail -1
done
