#!/bin/bash
# ncu --set full of every hot-path kernel once (second, warm pass of scripts/ncu_workload.py) -> gpurun_out/<dir>/r2_all_kernels.ncu-rep
O=${1:-gpurun_out/r2ncu}; mkdir -p $O
N=$(python scripts/ncu_workload.py | awk '/^pass 0/{print $4}')   # launches of one pass = what to skip
echo "launches per pass: $N"
timeout 1500 ncu --set full --clock-control none -k regex:^k_ -s $N -c 200 -o $O/r2_all_kernels python scripts/ncu_workload.py > $O/ncu_all.log 2>&1
echo "ncu rc=$?"; tail -3 $O/ncu_all.log; ls -la $O
