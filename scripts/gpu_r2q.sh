#!/bin/bash
# round-2 GPU check Q (8 GPUs): BASELINE configs[3] exactly -- 512 pairwise refines, 64 per robot/GPU
N=8; O=gpurun_out/r2q; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --pairs 64 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "cfg4 rc=$?"; tail -c 300 $O/bench_cfg4.err
python - <<EOF
import json
d=json.loads(open('$O/bench_cfg4.json').read().strip().splitlines()[-1])
print("cfg4", round(d["value"],1), round(d["ms_per_step"],3), d.get("parity"), d["clocks"]["per_rank"]["ms_per_step"], "e2e", round(d["e2e"]["value"],1))
EOF
