#!/bin/bash
# round-2 GPU check W: NN-1 tile scan on the packed-FP32 pipe (FADD2/FMUL2/FFMA2, four points per step) against the scalar scan
O=gpurun_out/r2w; mkdir -p $O
timeout 900 python -m pytest tests/test_gicp_gpu.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest.log
for v in default x2off; do
  if [ $v = default ]; then unset MRSLAM_LIB; else export MRSLAM_LIB=$PWD/build/variants/libmrslam_$v.so; fi
  echo "== $v"; timeout 300 python scripts/prof_gicp.py 8 2 2>&1 | grep -E "per-kernel|nn1 pair evals" | tail -2
done
unset MRSLAM_LIB
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench rc=$?"
python - <<EOF
import json
d=json.loads(open('$O/bench_cfg2.json').read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
print({k: v["ms_per_step"] for k, v in d["roofline"]["groups"].items()})
EOF
# source-level counters of the two search kernels (a seeded NN-1 pass: skip the first two launches)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_nn1_search -s 2 -c 1 -o $O/r2w_nn1 python scripts/prof_gicp.py 2 1 > $O/ncu_nn1.log 2>&1; echo "ncu nn1 rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:^k_knn_search -c 1 -o $O/r2w_knn python scripts/prof_gicp.py 2 1 > $O/ncu_knn.log 2>&1; echo "ncu knn rc=$?"
ls -la $O
