#!/bin/bash
# round-2 GPU check S: map_lowest atomic A/B (hardware min/max vs CAS loop) on the two elevation benches, after the reverts
O=gpurun_out/r2s; mkdir -p $O
for v in default elevcas; do
  if [ $v = default ]; then unset MRSLAM_LIB; else export MRSLAM_LIB=$PWD/build/variants/libmrslam_$v.so; fi
  timeout 400 python bench.py --workload tiled --steps 5 --warmup 3 > $O/tiled_$v.json 2> $O/tiled_$v.err
  timeout 400 python bench.py --steps 8 --warmup 3 > $O/cfg2_$v.json 2> $O/cfg2_$v.err
  python - <<EOF
import json
d=json.loads(open('$O/tiled_$v.json').read().strip().splitlines()[-1]); print('$v tiled', round(d["ms_per_step"],3), d["other_near_filter_setting"]["ms_per_step"])
d=json.loads(open('$O/cfg2_$v.json').read().strip().splitlines()[-1]); g=d["roofline"]["groups"]; print('$v cfg2', round(d["ms_per_step"],3), g["elev_process"]["ms_per_step"], g["elev_fuse"]["ms_per_step"], d["elevation"]["ms_per_step"])
EOF
done
unset MRSLAM_LIB
timeout 900 python -m pytest tests/test_elevation_gpu.py -q 2>&1 | tail -2
