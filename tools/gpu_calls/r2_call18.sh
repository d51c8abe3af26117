#!/bin/bash
# Round 2, call 18 (--gpus 2): soak test of the cross-GPU exchange (tools/stress_sharded.py).
set -u
O=gpurun_out/r2c18; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tools/stress_sharded.py --seconds 60 2> $O/err.txt | grep '^{' > $O/stress_sharded_2gpu.json; echo "rc=$?"
cat $O/stress_sharded_2gpu.json; tail -5 $O/err.txt
