#!/bin/bash
# Round 2, call 24 (1 GPU): fp8 storage with the integer decode (codes read as floats against an eigenvector pre-scaled by
# 2^120): parity tests, then timings at three sizes.
set -u
O=gpurun_out/r2c24; mkdir -p $O
timeout 600 python -m pytest tests/test_zzz_gpu_fp8_storage.py -m gpu -q -x > $O/pytest.txt 2>&1; tail -4 $O/pytest.txt
B="--no-cpu-baseline --no-sweep-table --no-e2e --no-north-star --steps 5"
for w in hilbert-8192 hilbert-16384 hilbert-32768 hilbert-65536; do
  timeout 300 python bench.py --workload $w --storage fp8 $B >> $O/storage.json 2>> $O/err.txt
done
python - $O/storage.json <<'PY'
import json,sys
for line in open(sys.argv[1]):
    d=json.loads(line)
    print(d["config"]["workload"], d["config"]["storage"], d["roofline"]["kernel"][:34], d["value"], d["roofline"]["frac"], d["us_per_round"], d["phase_us"], d["eigen_val"], d["rounds"])
PY
tail -5 $O/err.txt
