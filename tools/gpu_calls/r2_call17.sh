#!/bin/bash
# Round 2, call 17 (--gpus 2): the resident-e tail with this GPU's max word read together with the first batch of s
# (one L2 round trip after the barrier instead of two): parity, then timings at 1 and 2 GPUs.
set -u
O=gpurun_out/r2c17; mkdir -p $O
B="--no-cpu-baseline --no-e2e --no-north-star"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_bitexact.py tests/test_gpu_sharded.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --scale-base-dim 32768 $B > $O/bench_n1.json 2>> $O/err.txt
timeout 300 python bench.py --workload hilbert-16384 --steps 10 --no-sweep-table $B >> $O/more.json 2>> $O/err.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-north-star 2>>$O/err.txt | grep '^{' >> $O/more.json
timeout 300 $TR --nproc-per-node 2 --master-port 29603 bench.py --gpus 2 --workload hilbert-8192 --steps 10 --warmup 3 --no-e2e --no-north-star 2>>$O/err.txt | grep '^{' >> $O/more.json
python - $O/bench_n1.json $O/more.json <<'PY'
import json,sys
for f in sys.argv[1:]:
  for line in open(f):
    d=json.loads(line)
    print(d["config"]["workload"], d["n_gpus"], d["value"], d["roofline"]["frac"], d["us_per_round"], d["phase_us"], d["rounds"], (d.get("parity") or {}).get("bits_equal"))
    if d.get("hilbert_sweep"): print("  sweep", [(r["N"], r["ms_to_converge"], r["us_per_round"]) for r in d["hilbert_sweep"]])
    if d.get("strong_scaling_base"): print("  base", d["strong_scaling_base"]["value"], d["strong_scaling_base"]["ms_to_converge"])
PY
tail -3 $O/err.txt
