#!/bin/bash
# round-2 GPU check AB: pending host uploads unpacked by ONE launch per build (was: 64 event waits + 64 small kernels per step)
O=gpurun_out/r2ab; mkdir -p $O
timeout 900 python -m pytest tests/test_gicp_gpu.py tests/test_cpp_adapters.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
BENCH_E2E_DIAG=1 timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench rc=$?"; grep "e2e diag" $O/bench_cfg2.err
python - <<EOF
import json
d=json.loads(open('$O/bench_cfg2.json').read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],3), "single", round(d["e2e"]["single_batch"]["ms_per_step"],3), "compat", round(d["e2e"]["compat"]["value"],1), d["e2e"]["transforms_bitwise_equal_device_path"])
EOF
