#!/bin/bash
# Round 2, call 23 (1 GPU): fp8 storage on hardware for the first time (parity tests, then timings at three sizes next to
# bf16 and fp32), the ragged-dimension tests on the final tail, and one ncu capture of the scalar-unit resident-e kernel.
set -u
O=gpurun_out/r2c23; mkdir -p $O
timeout 600 python -m pytest tests/test_zzz_gpu_fp8_storage.py tests/test_zz_gpu_bitexact.py -m gpu -q -x > $O/pytest.txt 2>&1; tail -4 $O/pytest.txt
B="--no-cpu-baseline --no-sweep-table --no-e2e --no-north-star --steps 5"
for w in hilbert-8192 hilbert-32768; do
  for st in fp8 bf16 f32; do
    timeout 300 python bench.py --workload $w --storage $st $B >> $O/storage.json 2>> $O/err.txt
  done
done
timeout 300 python bench.py --workload hilbert-65536 --storage fp8 $B >> $O/storage.json 2>> $O/err.txt
timeout 300 python bench.py --workload hilbert-8191 $B >> $O/storage.json 2>> $O/err.txt
python - $O/storage.json <<'PY'
import json,sys
for line in open(sys.argv[1]):
    d=json.loads(line)
    print(d["config"]["workload"], d["config"]["storage"], d["roofline"]["kernel"][:34], d["value"], d["roofline"]["frac"], d["us_per_round"], d["phase_us"], d["eigen_val"], d["rounds"])
PY
python tools/profile_target.py hilbert-8191 3 > $O/plain.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:round_loop -s 1 -c 1 -o $O/ragged8191 python tools/profile_target.py hilbert-8191 3 > $O/ncu.log 2>&1
ncu -i $O/ragged8191.ncu-rep --page details > $O/ragged8191_details.txt 2>&1
ncu -i $O/ragged8191.ncu-rep --page raw --csv > $O/ragged8191_raw.csv 2>&1
ncu -i $O/ragged8191.ncu-rep --page source --csv > $O/ragged8191_source.csv 2>&1
rm -f $O/ragged8191.ncu-rep
cat $O/plain.log; tail -5 $O/err.txt
