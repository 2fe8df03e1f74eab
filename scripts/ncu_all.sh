#!/bin/bash
# ncu capture of every hot-path kernel once (second, warm pass of scripts/ncu_workload.py) -> <dir>/r2_all_kernels.ncu-rep
# Sections instead of --set full keep the report well under gpurun's 64 MiB return limit (59 kernels x 1.1 MB did not fit).
O=${1:-gpurun_out/r2ncu}; mkdir -p $O
N=$(python scripts/ncu_workload.py | awk '/^pass 0/{print $4}')   # launches of one pass = what to skip
echo "launches per pass: $N"
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --section SchedulerStats --section WarpStateStats --section ComputeWorkloadAnalysis --section InstructionStats"
timeout 1500 ncu $SEC --clock-control none -k regex:^k_ -s $N -c 200 -o $O/r2_all_kernels python scripts/ncu_workload.py > $O/ncu_all.log 2>&1
echo "ncu rc=$?"; tail -2 $O/ncu_all.log; ls -la $O
