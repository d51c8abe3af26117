#!/bin/bash
# Round 2, call 26 (1 GPU): what the driver does at round end, on the final tree -- the whole -m gpu suite, smoke(), the
# reference arm and the default bench line -- then the launch list and one ncu --set full capture of the default kernel
# on this build, and the last timings of the storage formats and the ragged dimensions.
set -u
O=gpurun_out/r2c26; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rs --durations=6 > $O/pytest_gpu_full.txt 2>&1; tail -12 $O/pytest_gpu_full.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_n1_reference_arm.json 2> $O/err.txt
timeout 900 python bench.py > $O/bench_n1.json 2>> $O/err.txt; echo "bench rc=$?"
python - $O/bench_n1.json $O/bench_n1_reference_arm.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read()); r=json.loads(open(sys.argv[2]).read())
print(d["value"], d["roofline"]["frac"], d["us_per_round"], d["ms_to_converge"], d["phase_us"], d["parity"]["bits_equal"], d["e2e"]["ms_per_step"], d["e2e_pageable"]["ms_per_step"], d["clocks"])
for x in d["north_star"]: print(x["workload"], x["value"], x["frac"], x["parity"]["bits_equal"])
print("reference arm", r["value"], r["ms_per_step"], r["cpu_baseline"]["cores"], "e2e ratio", round(d["e2e"]["value"]/r["e2e"]["value"],1), "device ratio", round(d["value"]/r["value"],1))
PY
B="--no-cpu-baseline --no-sweep-table --no-e2e --no-north-star --steps 5"
for w in hilbert-8192 hilbert-32768; do
  for st in fp8 bf16; do
    timeout 300 python bench.py --workload $w --storage $st $B >> $O/storage.json 2>> $O/err.txt
  done
done
timeout 300 python bench.py --workload hilbert-65536 --storage fp8 $B >> $O/storage.json 2>> $O/err.txt
for w in hilbert-8191 hilbert-16383 hilbert-32767; do
  timeout 300 python bench.py --workload $w $B >> $O/storage.json 2>> $O/err.txt
done
python - $O/storage.json <<'PY'
import json,sys
for line in open(sys.argv[1]):
    d=json.loads(line)
    print(d["config"]["workload"], d["config"]["storage"], d["roofline"]["kernel"][:34], d["value"], d["roofline"]["frac"], d["us_per_round"], d["phase_us"], d["eigen_val"], d["rounds"])
PY
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-north-star --no-sweep-table > $O/plain_bench.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-north-star --no-sweep-table > $O/ncu_launches.log 2>&1
python tools/profile_target.py hilbert-8192 3 > $O/plain.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:round_loop -s 1 -c 1 -o $O/h8192 python tools/profile_target.py hilbert-8192 3 > $O/ncu.log 2>&1
ncu -i $O/h8192.ncu-rep --page details > $O/h8192_details.txt 2>&1
ncu -i $O/h8192.ncu-rep --page raw --csv > $O/h8192_raw.csv 2>&1
rm -f $O/h8192.ncu-rep
cat $O/plain.log; tail -5 $O/err.txt
