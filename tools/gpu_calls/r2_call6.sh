#!/bin/bash
# Round 2, call 6 (1 GPU): end game v3 (pipelined shares through shared memory): off / 2 / 4 / 8 units per CTA.
set -u
O=gpurun_out/r2c6; mkdir -p $O
B="--no-cpu-baseline --no-sweep-table --no-e2e --no-north-star"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_bitexact.py -m gpu -q -x > $O/pytest.txt 2>&1
tail -3 $O/pytest.txt
for w in hilbert-8192 hilbert-16384 hilbert-32768; do
  timeout 300 python bench.py --workload $w --sweep 17 --steps 10 $B >> $O/endgame.json 2>> $O/err.txt
  for eg in 1 2 4 8; do
    ST_ENDGAME=$eg timeout 300 python bench.py --workload $w --sweep 1 --steps 10 $B >> $O/endgame.json 2>> $O/err.txt
  done
done
python - $O/endgame.json <<'PY'
import json,sys
for line in open(sys.argv[1]):
    d=json.loads(line)
    print(d["config"]["workload"], "sweep",d["config"]["sweep"], d["value"], d["roofline"]["frac"], d["us_per_round"], d["phase_us"], d["eigen_val"], d["rounds"], d.get("parity",{}) and d["parity"].get("bits_equal"))
PY
tail -3 $O/err.txt
