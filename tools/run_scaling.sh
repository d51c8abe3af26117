#!/bin/bash
# Multi-GPU parity + scaling lines; run under: gpurun --gpus 8 -- bash tools/run_scaling.sh
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -6
for n in 2 4 8; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29600 + n)) bench.py --gpus $n --steps 5 --warmup 3 2>gpurun_out/scale_err_n$n.txt | grep '^{' > gpurun_out/scale_hilbert32768_n$n.json
  cut -c1-330 gpurun_out/scale_hilbert32768_n$n.json
done
timeout 300 python bench.py --workload hilbert-32768 --steps 5 --no-cpu-baseline --no-sweep-table 2>/dev/null | grep '^{' > gpurun_out/scale_hilbert32768_n1.json
cut -c1-400 gpurun_out/scale_hilbert32768_n1.json
timeout 600 $TR --nproc-per-node 8 --master-port 29650 bench.py --gpus 8 --workload hilbert-131072 --steps 3 --warmup 3 --no-e2e 2>/dev/null | grep '^{' > gpurun_out/scale_hilbert131072_n8.json
cut -c1-1200 gpurun_out/scale_hilbert131072_n8.json
timeout 600 $TR --nproc-per-node 8 --master-port 29651 bench.py --gpus 8 --workload uniform-131072 --steps 1 --warmup 3 --no-e2e 2>/dev/null | grep '^{' > gpurun_out/scale_uniform131072_n8.json
cut -c1-1200 gpurun_out/scale_uniform131072_n8.json
timeout 600 $TR --nproc-per-node 8 --master-port 29652 bench.py --gpus 8 --workload uniform-65536 --steps 1 --warmup 3 --no-e2e 2>/dev/null | grep '^{' > gpurun_out/scale_uniform65536_n8.json
cut -c1-1200 gpurun_out/scale_uniform65536_n8.json
# in-process device group behind the drop-in handle (written without hardware): tests first, then one GPU vs all GPUs, e2e
timeout 600 python -m pytest tests/test_zzzz_gpu_group.py tests/test_gpu_sharded.py -q -k "group or in_process or missing_peer" > gpurun_out/group_tests.txt 2>&1
tail -3 gpurun_out/group_tests.txt
for n in 8192 32768; do
  timeout 600 python tools/bench_group.py --dim $n >> gpurun_out/bench_group.json 2>> gpurun_out/bench_group.err
done
cat gpurun_out/bench_group.json
# L2 prefetch variant on the weakest scaling point (8 GPUs x Hilbert 32768: barrier 16.6 us + tail 7.5 us of a 93 us round)
for k in 0 22 23 24; do
  timeout 300 $TR --nproc-per-node 8 --master-port 2966$((k % 10)) bench.py --gpus 8 --workload hilbert-32768 --kernel $k --steps 10 --warmup 3 --no-e2e 2>/dev/null | grep '^{' >> gpurun_out/scale_hilbert32768_n8_l2_prefetch.json
done
cut -c1-400 gpurun_out/scale_hilbert32768_n8_l2_prefetch.json
