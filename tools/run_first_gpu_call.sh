#!/bin/bash
# First GPU call after a stretch of CPU-only work: everything that was written without hardware,
# in the order that gives the most information if the call is cut short.  One B200:
#
#   gpurun --timeout 1500 -- 'bash tools/run_first_gpu_call.sh'
#
# Writes into gpurun_out/ (scratch); copy what should be judged into profiles/.
set -u
mkdir -p gpurun_out
O=gpurun_out
{
  echo "== smoke"
  timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
  echo "== bit-exact + relative-stop tests (new)"
  timeout 900 python -m pytest tests/test_zz_gpu_bitexact.py -q -x 2>&1 | tail -15
  echo "== bf16 storage tests (new; passed on hardware at the end of round 1)"
  timeout 900 python -m pytest tests/test_zzz_gpu_bf16_storage.py -q -rxX 2>&1 | tail -30
  echo "== full gpu suite"
  timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8
} > $O/first_call_tests.txt 2>&1
timeout 600 python bench.py > $O/first_call_bench_n1.json 2> $O/first_call_bench_n1.err
timeout 600 python tools/bench_kernels.py --json $O/first_call_bench_kernels.json > $O/first_call_bench_kernels.txt 2>&1
# relative stop on BASELINE config 4's generator at a single-GPU size: converges instead of running 1000 rounds
timeout 600 python bench.py --workload uniform-32768 --stop relative --eps 1e-6 --steps 5 --no-e2e --no-cpu-baseline \
  --no-sweep-table > $O/first_call_bench_uniform32768_relative.json 2>> $O/first_call_bench_n1.err
# bf16 storage: half the bytes per round; 8192^2 becomes L2-sized (128 MiB), 32768^2 is 2 GiB
for w in hilbert-8192 hilbert-32768; do
  timeout 600 python bench.py --workload $w --storage bf16 --steps 10 --no-cpu-baseline --no-sweep-table \
    > $O/first_call_bench_${w}_bf16.json 2>> $O/first_call_bench_n1.err
done
# fp64 accumulation: same bytes, two conversions and one DFMA per element -- does it still sit under the stream?
timeout 600 python bench.py --workload hilbert-8192 --accumulate f64 --steps 10 --no-e2e --no-cpu-baseline --no-sweep-table \
  > $O/first_call_bench_hilbert-8192_acc64.json 2>> $O/first_call_bench_n1.err
# streamed solve (written without hardware): half / 90 % of a 4 GiB matrix cached, pinned and pageable source
timeout 300 python -m pytest tests/test_zzzz_gpu_streamed.py -q > $O/first_call_streamed_tests.txt 2>&1
for c in 0.5 0.9; do
  timeout 600 python tools/bench_streamed.py --dim 32768 --cached $c >> $O/first_call_bench_streamed.json 2>> $O/first_call_bench_n1.err
done
timeout 600 python tools/bench_streamed.py --dim 32768 --cached 0.5 --pinned 0 >> $O/first_call_bench_streamed.json 2>> $O/first_call_bench_n1.err
# pageable host matrix through max_eigen_value: the driver's staging vs ST_UPLOAD_THREADS (opt-in until measured)
timeout 600 python tools/bench_upload.py --dim 8192 > $O/first_call_bench_upload.json 2>> $O/first_call_bench_n1.err
# zero-code experiment for the Hilbert-8192 last wave: 8192 units over 148 x W warps -- W = 14 gives 3.95
# units per warp (a nearly full last wave) against 3.46 at W = 16
for t in 384 416 448 480 512; do
  timeout 300 python bench.py --threads $t --steps 20 --no-e2e --no-cpu-baseline --no-sweep-table \
    >> $O/first_call_bench_threads_sweep.json 2>> $O/first_call_bench_n1.err
done
# L2 prefetch across the round barrier (kernels 21-23) against the default configuration 13: one GPU and sizes where
# barrier + tail are a visible share of the round
KERNELS=13,21,22,23 timeout 600 python tools/sweep_kernels.py 4096 8192 16384 32768 > $O/first_call_sweep_l2_prefetch.txt 2>&1
# ... and prefetch of the next unit during the pass (24-26) under static scheduling, against 13 static and 13 dynamic
KERNELS=13,24,25,26 SWEEPS=1,3 timeout 600 python tools/sweep_kernels.py 8192 32768 >> $O/first_call_sweep_l2_prefetch.txt 2>&1
tail -5 $O/first_call_tests.txt
