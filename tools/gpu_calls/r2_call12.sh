#!/bin/bash
# Round 2, call 12 (--gpus 8): final build -- multi-GPU parity tests, the scaling lines at N = 8 / 4 / 2 (+ the one-GPU base on
# the same box) with north-star side records and per-rank phase splits, BASELINE config 4 to its cap, the H2D ceiling of the
# box, the device group behind the drop-in handle.
set -u
O=gpurun_out/r2c12; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_zzzz_gpu_group.py tests/test_reference_wrapper_dropin.py -m gpu -q -rs > $O/multigpu_pytest_8gpu.txt 2>&1
tail -4 $O/multigpu_pytest_8gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
  timeout 600 $TR --nproc-per-node $n --master-port $((29600 + n)) bench.py --gpus $n --steps 5 --warmup 3 2>>$O/err.txt | grep '^{' > $O/bench_n$n.json; echo "n$n rc=$?"
done
timeout 300 python bench.py --no-cpu-baseline --no-north-star > $O/bench_n1_same_box.json 2>>$O/err.txt
timeout 600 $TR --nproc-per-node 8 --master-port 29618 bench.py --gpus 8 --workload uniform-65536 --steps 1 --warmup 3 --no-e2e 2>>$O/err.txt | grep '^{' > $O/bench_n8_uniform65536.json
timeout 600 $TR --nproc-per-node 8 --master-port 29628 bench.py --gpus 8 --workload hilbert-65536 --steps 3 --warmup 3 --no-e2e 2>>$O/err.txt | grep '^{' > $O/bench_n8_hilbert65536.json
timeout 300 python tools/h2d_probe.py > $O/h2d_probe.json 2>>$O/err.txt; cat $O/h2d_probe.json
for n in 8192 32768; do
  timeout 300 python tools/bench_group.py --dim $n >> $O/bench_group.json 2>> $O/err.txt
  timeout 300 python tools/bench_group.py --dim $n --pinned 0 >> $O/bench_group.json 2>> $O/err.txt
done
cat $O/bench_group.json
python - $O/bench_n8.json $O/bench_n4.json $O/bench_n2.json $O/bench_n1_same_box.json $O/bench_n8_uniform65536.json $O/bench_n8_hilbert65536.json <<'PY'
import json,sys
for f in sys.argv[1:]:
  for line in open(f):
    d=json.loads(line)
    print(d["config"]["workload"], d["n_gpus"], d["value"], d["roofline"]["frac"], d["us_per_round"], d["phase_us"], d["eigen_val"], d["rounds"], (d.get("parity") or {}).get("bits_equal"), d["clocks"])
    if d.get("phase_us_by_rank"): print("   ranks", [(r["pass_us"], r["barrier_us"], r["tail_us"]) for r in d["phase_us_by_rank"]])
    for k in ("e2e","e2e_pageable"):
        if d.get(k): print("  ",k,d[k]["ms_per_step"],d[k]["value"])
    if d.get("strong_scaling_base"): print("   base", d["strong_scaling_base"])
    for r in d.get("north_star") or []:
        print("   NS", r["workload"], r["value"], r["frac"], r["us_per_round"], r["phase_us"], r["rounds"], r["parity"]["bits_equal"], r["kernel"], r["clocks"])
        if r.get("phase_us_by_rank"): print("      ranks", [(x["pass_us"], x["barrier_us"], x["tail_us"]) for x in r["phase_us_by_rank"]])
PY
tail -5 $O/err.txt
