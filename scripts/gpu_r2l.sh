#!/bin/bash
# round-2 GPU check L (4 GPUs): where do the 3-4 ms per step go at N=4?  (kernel times are unchanged vs N=1)
N=4; O=gpurun_out/r2l; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
run() { name=$1; shift; env "$@" timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/$name.json 2> $O/$name.err; python - <<EOF
import json
d=json.loads(open('$O/$name.json').read().strip().splitlines()[-1])
print('$name', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['per_rank']['ms_per_step'], d['cpu_affinity'])
EOF
}
run default A=1
run sample_every_100 BENCH_SAMPLE_EVERY=100
run nobind BENCH_NO_BIND=1
run nobind_nosample BENCH_NO_BIND=1 BENCH_SAMPLE_EVERY=100
nproc; lscpu | grep -E "NUMA|Socket|Model name" | head; nvidia-smi topo -m | head -14
