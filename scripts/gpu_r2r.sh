#!/bin/bash
# round-2 GPU check R: parity suite + benches after the map_lowest atomic, the walk speculation and the async-allocation ordering
O=gpurun_out/r2r; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
timeout 400 python bench.py --workload tiled --steps 5 --warmup 3 > $O/bench_tiled.json 2> $O/bench_tiled.err; echo "tiled rc=$?"
timeout 400 python bench.py --steps 10 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "cfg2 rc=$?"
timeout 300 python bench.py --workload ring --steps 10 --warmup 3 > $O/bench_ring.json 2> $O/bench_ring.err; echo "ring rc=$?"
python - <<EOF
import json
for w in ("tiled","cfg2","ring"):
    d=json.loads(open('$O/bench_%s.json'%w).read().strip().splitlines()[-1])
    print(w, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d.get("parity"))
    if w=="tiled": print("   ", d["other_near_filter_setting"])
    if w=="cfg2": print("   ", {k:v["ms_per_step"] for k,v in d["roofline"]["groups"].items()}, d["elevation"]["ms_per_step"], d["elevation"]["value"])
EOF
