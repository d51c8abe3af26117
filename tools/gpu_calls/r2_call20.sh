#!/bin/bash
# Round 2, call 20 (--gpus 2): the flat cross-GPU barrier (every CTA arrives at every GPU's counter itself) against the
# forwarding-flag protocol (--sweep 17), same build, same box: sharded / group tests, A/B at three sizes, a short soak.
# (Ran at commit d2515d7: the forwarding-flag protocol, its --sweep 17 switch and tools/ab_barrier.py were removed after this
# measurement; check that commit out to repeat it.)
set -u
O=gpurun_out/r2c20; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_zzzz_gpu_group.py -m gpu -q -rs -x > $O/multigpu_pytest.txt 2>&1
tail -6 $O/multigpu_pytest.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
B="--gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-sweep-table --no-e2e --no-north-star"
port=29610
for w in hilbert-32768 hilbert-16384 hilbert-8192; do
  for sw in 1 17 1 17; do
    port=$((port+1))
    timeout 200 $TR --master-port $port bench.py $B --workload $w --sweep $sw 2>>$O/err.txt | grep '^{' >> $O/ab.json
  done
done
timeout 120 $TR --master-port 29650 tools/stress_sharded.py --seconds 8 2>> $O/err.txt | grep '^{' > $O/stress.json
python - $O/ab.json <<'PY'
import json,sys
for line in open(sys.argv[1]):
    d=json.loads(line)
    print(d["config"]["workload"], "sweep", d["config"]["sweep"], d["value"], d["roofline"]["frac"], d["us_per_round"], d["phase_us"], d["rounds"], d["parity"]["bits_equal"])
PY
cut -c1-600 $O/stress.json; tail -5 $O/err.txt
