#!/bin/bash
# Round 2, call 29 (--gpus 8): the final tree on 8 GPUs -- the sharded parity worker (bitwise == 1 GPU at G = 8) and the
# Hilbert 32768 bench line with every rank's phase split (no north-star side records: the budget's last minutes).
set -u
O=gpurun_out/r2c29; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 120 $TR --nproc-per-node 8 --master-port 29608 bench.py --gpus 8 --steps 5 --warmup 3 --no-north-star 2>>$O/err.txt | grep '^{' > $O/bench_n8.json; echo "bench rc=$?"
timeout 150 $TR --nproc-per-node 8 --master-port 29556 tests/sharded_worker.py > $O/sharded_worker_8gpu.txt 2>&1; echo "worker rc=$?"; grep -c SHARDED_OK $O/sharded_worker_8gpu.txt
python - $O/bench_n8.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print(d["config"]["workload"], d["n_gpus"], d["value"], d["roofline"]["frac"], d["us_per_round"], d["phase_us"], d["rounds"], d["parity"]["bits_equal"], d["clocks"])
if d.get("phase_us_by_rank"): print("   ranks", [(r["pass_us"], r["barrier_us"], r["tail_us"]) for r in d["phase_us_by_rank"]])
for k in ("e2e","e2e_pageable"):
    if d.get(k): print("  ",k,d[k]["ms_per_step"])
PY
tail -3 $O/err.txt
