#!/bin/bash
# Round 2, call 14 (1 GPU): smoke() on the final build; ncu --set full of the resident-e kernel at Hilbert 32768 (HBM-bound case).
set -u
O=gpurun_out/r2c14; mkdir -p $O
timeout 300 python __graft_entry__.py --smoke > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
T="python tools/profile_target.py hilbert-32768 2"
$T > $O/prof_32768_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:round_loop -s 1 -c 1 -o $O/prof_32768 $T > $O/prof_32768_ncu.log 2>&1
ncu -i $O/prof_32768.ncu-rep --page raw --csv > $O/ncu_full_round_loop_hilbert32768_raw.csv 2>/dev/null
ncu -i $O/prof_32768.ncu-rep --page details > $O/ncu_full_round_loop_hilbert32768_details.txt 2>/dev/null
cat $O/prof_32768_plain.log
timeout 300 python tools/bench_table.py --max-log2 17 > $O/reference_format_table.txt 2>&1; cat $O/reference_format_table.txt
