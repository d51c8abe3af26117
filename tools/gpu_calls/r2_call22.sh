#!/bin/bash
# Round 2, call 22 (1 GPU): dim % 4 != 0 on the resident-e kernel's scalar-unit build (automatic) against the general loop
# (--kernel 1), bit-exact tests of the ragged dimensions, and the aligned neighbour for the ratio.
set -u
O=gpurun_out/r2c22; mkdir -p $O
B="--no-cpu-baseline --no-sweep-table --no-e2e --no-north-star --steps 5"
timeout 300 python -m pytest tests/test_zz_gpu_bitexact.py -m gpu -q -x 2>&1 | tail -3
for w in hilbert-8191 hilbert-8190 hilbert-16383 hilbert-32767 hilbert-4099; do
  timeout 300 python bench.py --workload $w $B >> $O/ragged.json 2>> $O/err.txt
  timeout 300 python bench.py --workload $w --kernel 1 $B >> $O/ragged.json 2>> $O/err.txt
done
timeout 300 python bench.py --workload hilbert-8192 $B >> $O/ragged.json 2>> $O/err.txt
python - $O/ragged.json <<'PY'
import json,sys
for line in open(sys.argv[1]):
    d=json.loads(line)
    print(d["config"]["workload"], "kernel", d["roofline"]["kernel"][:40], d["value"], d["roofline"]["frac"], d["us_per_round"], d["phase_us"], d["eigen_val"], d["rounds"])
PY
tail -5 $O/err.txt
