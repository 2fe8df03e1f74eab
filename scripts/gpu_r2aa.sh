#!/bin/bash
# round-2 GPU check AA: where the e2e step loses 2.7 ms against the device-resident one (host phase times, kernel sums under e2e)
O=gpurun_out/r2aa; mkdir -p $O
BENCH_E2E_DIAG=1 timeout 600 python bench.py --steps 8 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench rc=$?"; grep "e2e diag" $O/bench_cfg2.err
python - <<EOF
import json
d=json.loads(open('$O/bench_cfg2.json').read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],3), "single", round(d["e2e"]["single_batch"]["ms_per_step"],3))
print({k: v["ms_per_step"] for k, v in d["roofline"]["groups"].items()})
EOF
