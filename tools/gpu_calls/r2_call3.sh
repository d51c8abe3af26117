#!/bin/bash
# Round 2, call 3 (--gpus 2): the reworked round barrier (max carried by the barrier and the cross-GPU flags, every CTA
# polls the flags) and the single-pass vector tail, on hardware: parity first, then timings at 1 and 2 GPUs.
set -u
O=gpurun_out/r2c3; mkdir -p $O
B="--no-cpu-baseline --no-sweep-table"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_bitexact.py tests/test_gpu_sharded.py tests/test_zzzz_gpu_group.py -m gpu -q -x > $O/pytest.txt 2>&1
tail -4 $O/pytest.txt
timeout 300 python bench.py --steps 20 --no-e2e $B > $O/bench_n1_8192.json 2>> $O/err.txt
timeout 300 python bench.py --workload hilbert-32768 --steps 5 --no-e2e $B > $O/bench_n1_32768.json 2>> $O/err.txt
timeout 300 python bench.py --workload hilbert-16384 --steps 10 --no-e2e $B > $O/bench_n1_16384.json 2>> $O/err.txt
timeout 300 python bench.py --workload hilbert-4096 --steps 20 --no-e2e $B > $O/bench_n1_4096.json 2>> $O/err.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e 2>>$O/err.txt | grep '^{' > $O/scale_hilbert32768_n2.json
timeout 300 $TR --nproc-per-node 2 --master-port 29603 bench.py --gpus 2 --workload hilbert-16384 --steps 10 --warmup 3 --no-e2e 2>>$O/err.txt | grep '^{' > $O/scale_hilbert16384_n2.json
for f in $O/*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print(sys.argv[1], d["config"]["workload"], d["n_gpus"], d["value"], d["roofline"]["frac"], d["us_per_round"], d["phase_us"], d["eigen_val"], d["rounds"])
PY
done
tail -3 $O/err.txt
