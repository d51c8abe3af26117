#!/bin/bash
# Round 2, call 8 (--gpus 2): the full default bench line at N = 2 (north-star side records sharded, pageable e2e leg),
# and the scalar row reduction with 32 loads in flight per lane (dim % 4 != 0).
set -u
O=gpurun_out/r2c8; mkdir -p $O
B="--no-cpu-baseline --no-sweep-table --no-e2e"
for w in hilbert-8191 hilbert-8190 hilbert-16383; do
  timeout 300 python bench.py --workload $w --steps 5 $B >> $O/ragged.json 2>> $O/err.txt
done
timeout 300 python bench.py --workload hilbert-8192 --form 1 --steps 10 $B >> $O/ragged.json 2>> $O/err.txt
timeout 300 python -m pytest tests/test_zz_gpu_bitexact.py -m gpu -q -x 2>&1 | tail -2
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --steps 5 --warmup 3 2>>$O/err.txt | grep '^{' > $O/bench_n2.json; echo "rc=$?"
python - $O/ragged.json $O/bench_n2.json <<'PY'
import json,sys
for f in sys.argv[1:]:
  for line in open(f):
    d=json.loads(line)
    print(d["config"]["workload"], d["n_gpus"], d["config"]["form"], d["value"], d["roofline"]["frac"], d["us_per_round"], d["phase_us"], d["eigen_val"], d["rounds"], d.get("parity"))
    for k in ("e2e","e2e_pageable"):
        if d.get(k): print("  ",k,d[k]["ms_per_step"],d[k]["value"])
    for r in d.get("north_star") or []:
        print("   NS", r["workload"], r["value"], r["frac"], r["us_per_round"], r["phase_us"], r["rounds"], r["parity"]["bits_equal"], r["clocks"])
PY
tail -5 $O/err.txt
