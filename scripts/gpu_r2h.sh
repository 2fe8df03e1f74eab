#!/bin/bash
# round-2 GPU check H: full parity suite after the k_cov change + bench lines + per-kernel ncu (two reports < 64 MiB together)
O=gpurun_out/r2h; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest.log
timeout 300 python scripts/prof_gicp.py 8 2 2>&1 | grep -E "per-kernel" | tail -1
timeout 400 python bench.py --steps 10 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "cfg2 rc=$?"; head -c 250 $O/bench_cfg2.json; echo; tail -3 $O/bench_cfg2.err
N=$(python scripts/ncu_workload.py | awk '/^pass 0/{print $4}'); echo "launches per pass: $N"
timeout 1500 ncu --set full --clock-control none -k regex:^k_ -s $N -c 200 -o $O/r2_all_kernels python scripts/ncu_workload.py > $O/ncu_all.log 2>&1; echo "ncu rc=$?"; tail -2 $O/ncu_all.log; ls -la $O
