#!/bin/bash
# round-2 GPU check E: full parity suite on the current tree (new elev_post, adapters, radon gap) + all-kernel ncu capture
O=gpurun_out/r2e; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -s > $O/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error|matcher gap|raytracing removed|relative L2" $O/pytest.log | tail -12; grep -B5 -A25 "Error\|FAILED\|assert" $O/pytest.log | head -80
bash scripts/ncu_all.sh $O
