#!/bin/bash
# round-2 GPU check C: where does k_knn_search_v2 spend its time?  launch list (v2 launch vs the fallback launch) + one full ncu capture
O=gpurun_out/r2c; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_knn -c 12 --csv --log-file $O/knn_launches.csv python scripts/prof_gicp.py 8 1 > $O/l.log 2>&1; echo "launch list rc=$?"; grep -v "^==" $O/knn_launches.csv | cut -d, -f5,13-15 | head -40
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_knn_search_v2 -c 1 -o $O/r2_knn_search_v2 python scripts/prof_gicp.py 8 1 > $O/ncu_knn.log 2>&1; echo "ncu rc=$?"; tail -3 $O/ncu_knn.log; ls -la $O
