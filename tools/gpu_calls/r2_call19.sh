#!/bin/bash
# Round 2, call 19 (--gpus 8): soak test of the cross-GPU exchange on 8 GPUs (tools/stress_sharded.py).
set -u
O=gpurun_out/r2c19; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29556 tools/stress_sharded.py --seconds 30 2> $O/err.txt | grep '^{' > $O/stress_sharded_8gpu.json; echo "rc=$?"
cat $O/stress_sharded_8gpu.json; tail -5 $O/err.txt
