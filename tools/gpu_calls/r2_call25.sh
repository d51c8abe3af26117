#!/bin/bash
# Round 2, call 25 (1 GPU): narrow storage on 4-element units (fp8: 32-bit loads, bf16: 64-bit loads -- conflict-free reads of
# the eigenvector from shared memory), and the scalar-unit ring of bulk copies for dim % 4 != 0 against the 4-byte loads
# (--sweep 17).  Parity tests first (the multi-GiB cases are left to the full suite at the end of the round).
# (Ran at commit 22a4650: the ring and its --sweep 17 switch were removed after this measurement.)
set -u
O=gpurun_out/r2c25; mkdir -p $O
timeout 600 python -m pytest tests/test_zzz_gpu_fp8_storage.py tests/test_zzz_gpu_bf16_storage.py tests/test_zz_gpu_bitexact.py -m gpu -q -x -k "not beyond" > $O/pytest.txt 2>&1; tail -4 $O/pytest.txt
B="--no-cpu-baseline --no-sweep-table --no-e2e --no-north-star --steps 5"
for w in hilbert-8192 hilbert-32768; do
  for st in fp8 bf16; do
    timeout 300 python bench.py --workload $w --storage $st $B >> $O/storage.json 2>> $O/err.txt
  done
done
timeout 300 python bench.py --workload hilbert-65536 --storage fp8 $B >> $O/storage.json 2>> $O/err.txt
for w in hilbert-8191 hilbert-8190 hilbert-12001 hilbert-4099; do
  timeout 300 python bench.py --workload $w $B >> $O/ragged.json 2>> $O/err.txt
  timeout 300 python bench.py --workload $w --sweep 17 $B >> $O/ragged.json 2>> $O/err.txt
done
python - $O/storage.json $O/ragged.json <<'PY'
import json,sys
for f in sys.argv[1:]:
  for line in open(f):
    d=json.loads(line)
    print(d["config"]["workload"], d["config"]["storage"], "sweep", d["config"]["sweep"], d["roofline"]["kernel"][:34], d["value"], d["roofline"]["frac"], d["us_per_round"], d["phase_us"], d["eigen_val"], d["rounds"])
PY
tail -5 $O/err.txt
