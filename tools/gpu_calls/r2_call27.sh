#!/bin/bash
# Round 2, call 27 (1 GPU): fp8 storage after the software-pipelined loop was reverted (call 26 measured it slower): parity
# tests of the storage formats and the last fp8 timings on the tree that ships.
set -u
O=gpurun_out/r2c27; mkdir -p $O
timeout 600 python -m pytest tests/test_zzz_gpu_fp8_storage.py tests/test_zzz_gpu_bf16_storage.py -m gpu -q -x -k "not beyond" > $O/pytest.txt 2>&1; tail -3 $O/pytest.txt
B="--no-cpu-baseline --no-sweep-table --no-e2e --no-north-star --steps 5"
for w in hilbert-8192 hilbert-16384 hilbert-32768 hilbert-65536; do
  timeout 300 python bench.py --workload $w --storage fp8 $B >> $O/storage.json 2>> $O/err.txt
done
timeout 300 python bench.py --workload hilbert-65536 --storage bf16 $B >> $O/storage.json 2>> $O/err.txt
python - $O/storage.json <<'PY'
import json,sys
for line in open(sys.argv[1]):
    d=json.loads(line)
    print(d["config"]["workload"], d["config"]["storage"], d["roofline"]["kernel"][:34], d["value"], d["roofline"]["frac"], d["us_per_round"], d["phase_us"], d["eigen_val"], d["rounds"])
PY
tail -5 $O/err.txt
