#!/bin/bash
# round-2 final GPU check T: whole parity suite, smoke(), the driver's default commands, refreshed per-kernel ncu report
O=gpurun_out/r2t; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench rc=$?"
timeout 300 python bench.py --workload ring --steps 10 --warmup 3 > $O/bench_ring.json 2> $O/bench_ring.err
timeout 400 python bench.py --workload tiled --steps 5 --warmup 3 > $O/bench_tiled.json 2> $O/bench_tiled.err
python - <<EOF
import json
for w in ("cfg2","ring","tiled"):
    d=json.loads(open('$O/bench_%s.json'%w).read().strip().splitlines()[-1])
    print(w, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "roofline", d["roofline"]["bound"], round(d["roofline"]["frac"],3))
EOF
bash scripts/ncu_all.sh $O
