#!/bin/bash
# Round 2, call 16 (1 GPU): dry run of what the driver does at round end, on the final tree: the whole -m gpu suite,
# smoke(), the default bench line and the reference arm.
set -u
O=gpurun_out/r2c16; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu_full.txt 2>&1; tail -3 $O/pytest_gpu_full.txt
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
tail -3 $O/err.txt
