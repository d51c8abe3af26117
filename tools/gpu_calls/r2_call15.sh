#!/bin/bash
# Round 2, call 15 (1 GPU): scalar streaming loads that allocate in L1 (dim % 4 != 0).
set -u
O=gpurun_out/r2c15; mkdir -p $O
B="--no-cpu-baseline --no-sweep-table --no-e2e"
timeout 600 python -m pytest tests/test_zz_gpu_bitexact.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
for w in hilbert-8191 hilbert-8190 hilbert-16383 hilbert-40001; do
  timeout 300 python bench.py --workload $w --steps 5 $B >> $O/ragged.json 2>> $O/err.txt
done
python - $O/ragged.json <<'PY'
import json,sys
for line in open(sys.argv[1]):
    d=json.loads(line)
    print(d["config"]["workload"], d["value"], d["roofline"]["frac"], d["us_per_round"], d["phase_us"], d["rounds"])
PY
tail -3 $O/err.txt
