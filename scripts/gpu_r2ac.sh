#!/bin/bash
# round-2 GPU check AC: per-stage device times of the cloud build in the device-resident steps and in the e2e steps
O=gpurun_out/r2ac; mkdir -p $O
MRSLAM_BUILD_DIAG=1 BENCH_E2E_DIAG=1 timeout 600 python bench.py --steps 8 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench rc=$?"; grep "diag" $O/bench_cfg2.err
