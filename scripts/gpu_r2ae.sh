#!/bin/bash
# round-2 GPU check AE (profiling evidence only): launch list of the default bench command, warp-instruction counts of the two search kernels
O=gpurun_out/r2ae; mkdir -p $O
timeout 100 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:"k_nn1_search|k_knn_search" -c 40 --csv --log-file $O/search_inst.csv python scripts/prof_gicp.py 2 1 > $O/inst.log 2>&1; echo "inst rc=$?"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --kill yes --csv --log-file $O/launches_bench_r2.csv python bench.py --steps 2 --warmup 1 > $O/launches.log 2>&1; echo "launches rc=$?"
wc -l $O/*.csv
