#!/bin/bash
# round-2 GPU check M (1 GPU): step time of every rank's pair set run alone -- is the N=4 "efficiency" data imbalance?
O=gpurun_out/r2m; mkdir -p $O
for r in 0 1 2 3 4 5 6 7; do
  BENCH_DATA_RANK=$r timeout 400 python bench.py --steps 8 --warmup 3 > $O/rank$r.json 2> $O/rank$r.err
  python - <<EOF
import json
d=json.loads(open('$O/rank$r.json').read().strip().splitlines()[-1])
print('data of rank $r alone:', round(d['ms_per_step'],2), 'ms/step', round(d['value'],1), 'aligns/s; linearize/align', round(d['gicp']['mean_linearize_per_align'],2), 'error evals', round(d['gicp']['mean_error_evals_per_align'],2))
EOF
done
