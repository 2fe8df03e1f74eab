#!/bin/bash
# round-2 GPU check AD: parameter tables fetched by a kernel (no DMA-queue stall behind bulk uploads): parity suite, smoke, the
# three bench workloads
O=gpurun_out/r2ad; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
BENCH_E2E_DIAG=1 timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench rc=$?"; grep "e2e diag" $O/bench_cfg2.err
python - <<EOF
import json
d=json.loads(open('$O/bench_cfg2.json').read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],3), "single", round(d["e2e"]["single_batch"]["ms_per_step"],3), "compat", round(d["e2e"]["compat"]["value"],1), d["e2e"]["transforms_bitwise_equal_device_path"], "launches", d["gpu_launches"])
print({k: v["ms_per_step"] for k, v in d["roofline"]["groups"].items()})
EOF
timeout 300 python bench.py --workload ring --steps 10 --warmup 3 > $O/bench_ring.json 2> $O/bench_ring.err
timeout 400 python bench.py --workload tiled --steps 5 --warmup 3 > $O/bench_tiled.json 2> $O/bench_tiled.err
python - <<EOF
import json
for w in ("ring","tiled"):
    d=json.loads(open('$O/bench_%s.json'%w).read().strip().splitlines()[-1])
    print(w, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
EOF
