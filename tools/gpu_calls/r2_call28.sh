#!/bin/bash
# Round 2, call 28 (--gpus 2): the multi-GPU tests on the final tree (flat barrier; the sharded ragged / bf16 / fp8 cases are
# new) and the driver-style bench line at N = 2.
set -u
O=gpurun_out/r2c28; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_zzzz_gpu_group.py tests/test_reference_wrapper_dropin.py -m gpu -q -rs > $O/multigpu_pytest_2gpu.txt 2>&1
tail -12 $O/multigpu_pytest_2gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --steps 5 --warmup 3 2>>$O/err.txt | grep '^{' > $O/bench_n2.json; echo "rc=$?"
python - $O/bench_n2.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print(d["config"]["workload"], d["n_gpus"], d["value"], d["roofline"]["frac"], d["us_per_round"], d["phase_us"], d["rounds"], d["parity"]["bits_equal"], d["clocks"])
for k in ("e2e","e2e_pageable"):
    if d.get(k): print("  ",k,d[k]["ms_per_step"])
for r in d.get("north_star") or []:
    print("   NS", r["workload"], r["value"], r["frac"], r["us_per_round"], r["rounds"], r["parity"]["bits_equal"])
PY
tail -5 $O/err.txt
