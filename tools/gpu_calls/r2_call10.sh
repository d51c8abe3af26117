#!/bin/bash
# Round 2, call 10 (--gpus 2): the wide kernel (unit-scheduled, windowed eigenvector) against the general loop at N > 32768.
set -u
O=gpurun_out/r2c10; mkdir -p $O
B="--no-cpu-baseline --no-sweep-table --no-e2e"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_bitexact.py tests/test_zz_gpu_options_property.py tests/test_gpu_sharded.py -m gpu -q -x > $O/pytest.txt 2>&1
tail -3 $O/pytest.txt
for k in 0 1; do
  timeout 300 python bench.py --workload hilbert-65536 --kernel $k --steps 5 $B >> $O/wide.json 2>> $O/err.txt
  timeout 300 python bench.py --workload hilbert-131072 --kernel $k --steps 2 $B >> $O/wide.json 2>> $O/err.txt
  timeout 300 python bench.py --workload hilbert-40960 --kernel $k --steps 5 $B >> $O/wide.json 2>> $O/err.txt
done
timeout 300 python bench.py --workload hilbert-32768 --steps 5 $B >> $O/wide.json 2>> $O/err.txt
timeout 300 python bench.py --workload hilbert-8192 --steps 20 $B >> $O/wide.json 2>> $O/err.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for k in 0 1; do
  timeout 300 $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --workload hilbert-65536 --kernel $k --steps 5 --warmup 3 --no-e2e 2>>$O/err.txt | grep '^{' >> $O/wide_n2.json
done
timeout 300 $TR --nproc-per-node 2 --master-port 29603 bench.py --gpus 2 --workload uniform-65536 --max-iter 50 --steps 2 --warmup 3 --no-e2e 2>>$O/err.txt | grep '^{' >> $O/wide_n2.json
python - $O/wide.json $O/wide_n2.json <<'PY'
import json,sys
for f in sys.argv[1:]:
  for line in open(f):
    d=json.loads(line)
    print(d["config"]["workload"], d["n_gpus"], "k",d["config"]["kernel"], d["roofline"]["kernel"][:28], d["value"], d["roofline"]["frac"], d["us_per_round"], d["phase_us"], d["rounds"], (d.get("parity") or {}).get("bits_equal"))
PY
tail -3 $O/err.txt
