#!/bin/bash
# round-2 GPU check Y: kNN with two queries per warp (k <= 16) + uniform insertion-round counts; NN-1 walk switches; 5-pass cloud sort
O=gpurun_out/r2y; mkdir -p $O
timeout 900 python -m pytest tests/test_gicp_gpu.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
for v in default pairoff oldins u2 pf1 k64 drop12; do
  if [ $v = default ]; then unset MRSLAM_LIB; else export MRSLAM_LIB=$PWD/build/variants/libmrslam_$v.so; fi
  echo "== $v"; timeout 300 python scripts/prof_gicp.py 8 2 2>&1 | grep -E "per-kernel|knn pair evals" | tail -2
done
unset MRSLAM_LIB
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench rc=$?"
python - <<EOF
import json
d=json.loads(open('$O/bench_cfg2.json').read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
print({k: v["ms_per_step"] for k, v in d["roofline"]["groups"].items()})
EOF
timeout 600 ncu --set full --import-source on --clock-control none -k regex:^k_knn_search -c 1 -o $O/r2y_knn python scripts/prof_gicp.py 2 1 > $O/ncu_knn.log 2>&1; echo "ncu knn rc=$?"
