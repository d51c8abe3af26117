#!/bin/bash
# Round 2, call 11 (1 GPU): the whole -m gpu suite on the final build, the default bench line, the reference arm,
# the launch list and the ncu --set full captures of the two dominant kernels.
set -u
O=gpurun_out/r2c11; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu_full.txt 2>&1; tail -3 $O/pytest_gpu_full.txt
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_n1_reference_arm.json 2>> $O/bench_n1.err
L="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-north-star --no-sweep-table"
$L > $O/launches_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_bench_steps2.csv $L > $O/launches_ncu.log 2>&1
T="python tools/profile_target.py hilbert-8192 3"
$T > $O/prof_8192_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:round_loop -s 1 -c 1 -o $O/prof_8192 $T > $O/prof_8192_ncu.log 2>&1
T="python tools/profile_target.py hilbert-65536 2"
$T > $O/prof_65536_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:round_loop -s 1 -c 1 -o $O/prof_65536 $T > $O/prof_65536_ncu.log 2>&1
for n in 8192 65536; do
  ncu -i $O/prof_$n.ncu-rep --page raw --csv > $O/ncu_full_round_loop_hilbert${n}_raw.csv 2>/dev/null
  ncu -i $O/prof_$n.ncu-rep --page details > $O/ncu_full_round_loop_hilbert${n}_details.txt 2>/dev/null
done
cat $O/prof_8192_plain.log $O/prof_65536_plain.log; ls -la $O | head -30
python - $O/bench_n1.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print(d["value"], d["roofline"], d["us_per_round"], d["phase_us"], d["parity"], d["e2e"], d["e2e_pageable"], d["cpu_baseline"]["value"], d["clocks"])
print(d["hilbert_sweep"]); print(d["strong_scaling_base"])
for r in d["north_star"]: print(r["workload"], r["value"], r["frac"], r["us_per_round"], r["phase_us"], r["parity"]["bits_equal"], r["kernel"])
PY
