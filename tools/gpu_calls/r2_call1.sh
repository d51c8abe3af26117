#!/bin/bash
# Round 2, first GPU call (one B200): the measurement backlog of round 1, ordered so that a cut-short call
# still leaves the most useful numbers.   gpurun --timeout 1380 -- 'bash tools/gpu_calls/r2_call1.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c1
mkdir -p $O
B="--no-cpu-baseline --no-sweep-table"
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
timeout 300 python __graft_entry__.py --smoke > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
# kernel variants: default 13 against the L2-prefetch variants, one GPU
KERNELS=13,21,22,23 timeout 400 python tools/sweep_kernels.py 8192 32768 > $O/sweep_l2_prefetch.txt 2>&1
KERNELS=13,24,25,26 SWEEPS=1,3 timeout 400 python tools/sweep_kernels.py 8192 32768 >> $O/sweep_l2_prefetch.txt 2>&1
for t in 384 448 480 512; do
  timeout 200 python bench.py --threads $t --steps 20 --no-e2e $B >> $O/threads_sweep.json 2>> $O/err.txt
done
# general loop at the sizes it serves: N > 32768 and ragged dims (VEC=1)
for w in hilbert-65536 hilbert-8191 hilbert-8190 hilbert-8188; do
  timeout 300 python bench.py --workload $w --steps 5 --no-e2e $B >> $O/general_loop.json 2>> $O/err.txt
done
timeout 300 python bench.py --workload hilbert-32768 --kernel 1 --steps 5 --no-e2e $B >> $O/general_loop.json 2>> $O/err.txt
# extensions: relative stop, bf16 storage, fp64 accumulation
timeout 300 python bench.py --workload uniform-32768 --stop relative --eps 1e-6 --steps 5 --no-e2e $B > $O/uniform32768_relative.json 2>> $O/err.txt
for w in hilbert-8192 hilbert-32768; do
  timeout 300 python bench.py --workload $w --storage bf16 --steps 10 $B >> $O/bf16.json 2>> $O/err.txt
done
timeout 300 python bench.py --workload hilbert-8192 --accumulate f64 --steps 10 --no-e2e $B > $O/acc64.json 2>> $O/err.txt
timeout 300 python bench.py --workload hilbert-32768 --accumulate f64 --steps 5 --no-e2e $B >> $O/acc64.json 2>> $O/err.txt
# per-kernel tables (reference benchmarks/similarity_transform.md counterpart)
timeout 500 python tools/bench_kernels.py --json $O/bench_kernels.json > $O/bench_kernels.txt 2>&1
# pageable host matrix through max_eigen_value
timeout 400 python tools/bench_upload.py --dim 8192 > $O/bench_upload.json 2>> $O/err.txt
# streamed solve
for c in 0.5 0.9; do
  timeout 400 python tools/bench_streamed.py --dim 32768 --cached $c >> $O/bench_streamed.json 2>> $O/err.txt
done
timeout 400 python tools/bench_streamed.py --dim 32768 --cached 0.5 --pinned 0 >> $O/bench_streamed.json 2>> $O/err.txt
timeout 300 python -m pytest tests/test_zzzz_gpu_streamed.py tests/test_zzzzz_gpu_l2_prefetch.py -q > $O/new_tests.txt 2>&1
tail -3 $O/new_tests.txt
ls -la $O
