#!/bin/bash
# round-2 GPU check D (N GPUs): multi-GPU legs of the three workloads (neighbour exchange + bitwise parity, tiled grid, RING shards)
N=${1:-2}; O=gpurun_out/r2d_n$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "cfg2 rc=$?"; tail -c 1500 $O/bench_cfg2.json | head -c 600; echo; tail -5 $O/bench_cfg2.err
timeout 600 $TR bench.py --gpus $N --workload tiled --steps 5 --warmup 3 > $O/bench_tiled.json 2> $O/bench_tiled.err; echo "tiled rc=$?"; head -c 300 $O/bench_tiled.json; echo; tail -5 $O/bench_tiled.err
timeout 600 $TR bench.py --gpus $N --workload ring --steps 10 --warmup 3 > $O/bench_ring.json 2> $O/bench_ring.err; echo "ring rc=$?"; head -c 300 $O/bench_ring.json; echo; tail -5 $O/bench_ring.err
