#!/bin/bash
# round-2 GPU check V: after the launch-bounds fix (NN-1 back to 56 registers, k_cov capped at 128): parity suite, smoke, per-kernel times, default bench
O=gpurun_out/r2v; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python scripts/prof_gicp.py 8 2 2>&1 | grep -E "per-kernel" | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench rc=$?"
python - <<EOF
import json
d=json.loads(open('$O/bench_cfg2.json').read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "roofline", d["roofline"]["bound"], round(d["roofline"]["frac"],3))
print(json.dumps(d["roofline"].get("groups", d["roofline"]))[:1500])
EOF
