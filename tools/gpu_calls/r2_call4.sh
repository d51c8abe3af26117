#!/bin/bash
# Round 2, call 4 (--gpus 2): end-game shares (the last units of a round handed out as 8 accumulator shares) on / off.
set -u
O=gpurun_out/r2c5; mkdir -p $O
B="--no-cpu-baseline --no-sweep-table --no-e2e"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_bitexact.py tests/test_gpu_sharded.py -m gpu -q -x > $O/pytest.txt 2>&1
tail -3 $O/pytest.txt
for w in hilbert-8192 hilbert-16384 hilbert-32768; do
  for sw in 1 17; do
    timeout 300 python bench.py --workload $w --sweep $sw --steps 10 $B >> $O/endgame_n1.json 2>> $O/err.txt
  done
done
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for w in hilbert-32768 hilbert-16384; do
  for sw in 1 17; do
    timeout 300 $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --workload $w --sweep $sw --steps 10 --warmup 3 --no-e2e 2>>$O/err.txt | grep '^{' >> $O/endgame_n2.json
  done
done
for f in $O/endgame_n1.json $O/endgame_n2.json; do python - "$f" <<'PY'
import json,sys
for line in open(sys.argv[1]):
    d=json.loads(line)
    print(d["config"]["workload"], "gpus",d["n_gpus"], "sweep",d["config"]["sweep"], d["value"], d["roofline"]["frac"], d["us_per_round"], d["phase_us"], d["eigen_val"], d["rounds"])
PY
done
tail -3 $O/err.txt
