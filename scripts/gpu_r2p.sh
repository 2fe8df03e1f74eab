#!/bin/bash
# round-2 GPU check P: linearize occupancy variants, e2e upload order, --pairs 64, parity of the bev/keyframe sort change
O=gpurun_out/r2p; mkdir -p $O
timeout 900 python -m pytest tests/test_bev_gpu.py tests/test_gicp_gpu.py -q 2>&1 | tail -2
for v in default lin6 lin8; do
  if [ $v = default ]; then unset MRSLAM_LIB; else export MRSLAM_LIB=$PWD/build/variants/libmrslam_$v.so; fi
  echo "== $v"; timeout 300 python scripts/prof_gicp.py 8 2 2>&1 | grep -E "per-kernel" | tail -1
done
unset MRSLAM_LIB
timeout 400 python bench.py --steps 10 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "cfg2 rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --pairs 64 > $O/bench_cfg2_p64.json 2> $O/bench_cfg2_p64.err; echo "cfg2 p64 rc=$?"
python - <<EOF
import json
for f in ("bench_cfg2","bench_cfg2_p64"):
    d=json.loads(open('$O/%s.json'%f).read().strip().splitlines()[-1])
    print(f, round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],2), d["e2e"]["transforms_bitwise_equal_device_path"])
EOF
