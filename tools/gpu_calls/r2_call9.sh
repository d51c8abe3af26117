#!/bin/bash
# Round 2, call 9 (--gpus 8): multi-GPU parity tests on hardware, the scaling lines with north-star side records,
# BASELINE configs 3-5 sharded over 8 GPUs, the device group behind the drop-in handle.
set -u
O=gpurun_out/r2c9; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" >> $O/topo.txt
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_zzzz_gpu_group.py tests/test_reference_wrapper_dropin.py -m gpu -q -rs > $O/multigpu_pytest.txt 2>&1
tail -6 $O/multigpu_pytest.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29608 bench.py --gpus 8 --steps 5 --warmup 3 2>>$O/err.txt | grep '^{' > $O/bench_n8.json; echo "n8 rc=$?"
timeout 600 $TR --nproc-per-node 4 --master-port 29604 bench.py --gpus 4 --steps 5 --warmup 3 2>>$O/err.txt | grep '^{' > $O/bench_n4.json; echo "n4 rc=$?"
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-north-star > $O/bench_n1_base.json 2>>$O/err.txt
timeout 600 $TR --nproc-per-node 8 --master-port 29618 bench.py --gpus 8 --workload uniform-65536 --steps 1 --warmup 3 --no-e2e 2>>$O/err.txt | grep '^{' > $O/bench_n8_uniform65536.json
timeout 600 $TR --nproc-per-node 8 --master-port 29628 bench.py --gpus 8 --workload hilbert-65536 --steps 3 --warmup 3 --no-e2e 2>>$O/err.txt | grep '^{' > $O/bench_n8_hilbert65536.json
for n in 8192 32768; do
  timeout 300 python tools/bench_group.py --dim $n >> $O/bench_group.json 2>> $O/err.txt
  timeout 300 python tools/bench_group.py --dim $n --pinned 0 >> $O/bench_group.json 2>> $O/err.txt
done
cat $O/bench_group.json
python - $O/bench_n8.json $O/bench_n4.json $O/bench_n1_base.json $O/bench_n8_uniform65536.json $O/bench_n8_hilbert65536.json <<'PY'
import json,sys
for f in sys.argv[1:]:
  for line in open(f):
    d=json.loads(line)
    print(d["config"]["workload"], d["n_gpus"], d["value"], d["roofline"]["frac"], d["us_per_round"], d["phase_us"], d["eigen_val"], d["rounds"], (d.get("parity") or {}).get("bits_equal"), d["clocks"])
    for k in ("e2e","e2e_pageable"):
        if d.get(k): print("  ",k,d[k]["ms_per_step"],d[k]["value"])
    if d.get("strong_scaling_base"): print("   base", d["strong_scaling_base"])
    for r in d.get("north_star") or []:
        print("   NS", r["workload"], r["value"], r["frac"], r["us_per_round"], r["phase_us"], r["rounds"], r["parity"]["bits_equal"], r["clocks"])
PY
tail -5 $O/err.txt
