#!/bin/bash
# round-2 GPU check I: k_cov (single gather pass, early-exit Newton), parity, per-kernel ncu sections, 8f-row timings, the three bench lines
O=gpurun_out/r2i; mkdir -p $O
timeout 900 python -m pytest tests/test_gicp_gpu.py tests/test_cpp_adapters.py -q -m gpu 2>&1 | tail -2
timeout 300 python scripts/prof_gicp.py 8 2 2>&1 | grep -E "per-kernel|tiles on the per-query" | tail -3
timeout 400 python bench.py --steps 10 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "cfg2 rc=$?"; head -c 250 $O/bench_cfg2.json; echo; tail -3 $O/bench_cfg2.err
timeout 300 python scripts/time_next_rows.py > $O/next_rows.json 2> $O/next_rows.err; echo "next rows rc=$?"; cat $O/next_rows.json; tail -3 $O/next_rows.err
bash scripts/ncu_all.sh $O
timeout 300 python bench.py --workload ring --steps 10 --warmup 3 > $O/bench_ring.json 2> $O/bench_ring.err; echo "ring rc=$?"
timeout 400 python bench.py --workload tiled --steps 5 --warmup 3 > $O/bench_tiled.json 2> $O/bench_tiled.err; echo "tiled rc=$?"
# NN-1: per-lane need masks + warp-per-pair visits (default) vs the round-1 shared lane-per-query walk (variant)
echo "== nn1 wpq (default)"; timeout 300 python scripts/prof_gicp.py 8 2 2>&1 | grep -E "per-kernel|nn1 pair evals" | tail -2
echo "== nn1 lpq (variant)"; MRSLAM_LIB=$PWD/build/variants/libmrslam_nn1lpq.so timeout 300 python scripts/prof_gicp.py 8 2 2>&1 | grep -E "per-kernel|nn1 pair evals" | tail -2
