#!/bin/bash
# round-2 GPU check Z: whole parity suite + smoke on the new defaults (two-query kNN, 5-pass cloud sort, deferred unpacking of host
# uploads), the three bench workloads (cfg2 now with the pipelined e2e region), NN-1 scan unroll variants
O=gpurun_out/r2z; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench rc=$?"; tail -3 $O/bench_cfg2.err
python - <<EOF
import json
d=json.loads(open('$O/bench_cfg2.json').read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],3), "single", d["e2e"]["single_batch"], "compat", round(d["e2e"]["compat"]["value"],1), "bitwise", d["e2e"]["transforms_bitwise_equal_device_path"])
print({k: v["ms_per_step"] for k, v in d["roofline"]["groups"].items()})
print(d["e2e"]["pipelined"])
EOF
for v in scan8 scan2; do
  export MRSLAM_LIB=$PWD/build/variants/libmrslam_$v.so
  echo "== $v"; timeout 300 python scripts/prof_gicp.py 8 2 2>&1 | grep -E "per-kernel" | tail -1
done
unset MRSLAM_LIB
timeout 300 python bench.py --workload ring --steps 10 --warmup 3 > $O/bench_ring.json 2> $O/bench_ring.err
timeout 400 python bench.py --workload tiled --steps 5 --warmup 3 > $O/bench_tiled.json 2> $O/bench_tiled.err
python - <<EOF
import json
for w in ("ring","tiled"):
    d=json.loads(open('$O/bench_%s.json'%w).read().strip().splitlines()[-1])
    print(w, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
EOF
