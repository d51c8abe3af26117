#!/bin/bash
# Round 2, call 2 (--gpus 2): everything multi-GPU that had only run on the emulation so far.
set -u
O=gpurun_out/r2c2; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_zzzz_gpu_group.py tests/test_reference_wrapper_dropin.py -m gpu -q -rs > $O/multigpu_pytest.txt 2>&1
tail -15 $O/multigpu_pytest.txt
for n in 8192 32768; do
  timeout 300 python tools/bench_group.py --dim $n >> $O/bench_group.json 2>> $O/err.txt
  timeout 300 python tools/bench_group.py --dim $n --pinned 0 >> $O/bench_group.json 2>> $O/err.txt
done
cat $O/bench_group.json
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --steps 5 --warmup 3 2>>$O/err.txt | grep '^{' > $O/scale_hilbert32768_n2.json
cut -c1-900 $O/scale_hilbert32768_n2.json
tail -5 $O/err.txt
