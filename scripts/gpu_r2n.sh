#!/bin/bash
# round-2 GPU check N (8 GPUs): the three workloads at N=8 (equal-work pair sets)
N=8; O=gpurun_out/r2d_n8; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus $N --steps 8 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "cfg2 rc=$?"; tail -c 400 $O/bench_cfg2.err
timeout 600 $TR bench.py --gpus $N --workload tiled --steps 5 --warmup 3 > $O/bench_tiled.json 2> $O/bench_tiled.err; echo "tiled rc=$?"; tail -c 400 $O/bench_tiled.err
timeout 600 $TR bench.py --gpus $N --workload ring --steps 8 --warmup 3 > $O/bench_ring.json 2> $O/bench_ring.err; echo "ring rc=$?"; tail -c 400 $O/bench_ring.err
python - <<EOF
import json
for w in ("cfg2","tiled","ring"):
    try:
        d=json.loads(open('$O/bench_%s.json'%w).read().strip().splitlines()[-1])
        print(w, round(d["value"],1), round(d["ms_per_step"],3), d.get("parity"), (d["clocks"].get("per_rank") or {}).get("ms_per_step"))
    except Exception as e: print(w, "ERR", e)
EOF
