#!/bin/bash
# Round 2, call 7 (1 GPU): general loop with a 32768-column staged window; the full default bench line (parity verdicts,
# north-star side records, pageable e2e leg) for the first time on hardware.
set -u
O=gpurun_out/r2c7; mkdir -p $O
B="--no-cpu-baseline --no-sweep-table --no-e2e"
timeout 900 python -m pytest tests/test_zz_gpu_bitexact.py -m gpu -q -x > $O/pytest.txt 2>&1
tail -3 $O/pytest.txt
for w in hilbert-65536 hilbert-8191; do
  timeout 300 python bench.py --workload $w --steps 5 $B >> $O/general_loop.json 2>> $O/err.txt
done
timeout 300 python bench.py --workload hilbert-32768 --kernel 1 --steps 5 $B >> $O/general_loop.json 2>> $O/err.txt
timeout 300 python bench.py --workload hilbert-8192 --form 1 --steps 10 $B >> $O/general_loop.json 2>> $O/err.txt
timeout 900 python bench.py > $O/bench_n1.json 2>> $O/err.txt; echo "bench rc=$?"
python - $O/general_loop.json $O/bench_n1.json <<'PY'
import json,sys
for f in sys.argv[1:]:
  for line in open(f):
    d=json.loads(line)
    print(d["config"]["workload"], d["config"]["form"], d["value"], d["roofline"]["frac"], d["us_per_round"], d["phase_us"], d["eigen_val"], d["rounds"], d.get("parity"))
    for k in ("e2e","e2e_pageable"):
        if d.get(k): print("  ",k,d[k]["ms_per_step"],d[k]["value"])
    for r in d.get("north_star") or []:
        print("   NS", r["workload"], r["value"], r["frac"], r["us_per_round"], r["phase_us"], r["rounds"], r["parity"], r["clocks"])
PY
tail -3 $O/err.txt
