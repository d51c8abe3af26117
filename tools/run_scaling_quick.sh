#!/bin/bash
# hilbert-32768 on N GPUs (N = $1), automatic warp count vs pinned 512 threads vs general kernel
set -u
n=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $n"
for extra in ${VARIANTS:-"" "--threads 512"}; do
  echo "== gpus=$n $extra"
  timeout 300 $TR --master-port $((29700 + RANDOM % 200)) bench.py --gpus $n --steps 5 --warmup 3 --no-e2e --workload ${WORKLOAD:-hilbert-32768} $extra 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','us_per_round','rounds')}, d['roofline']['frac'], d['roofline']['kernel'], d.get('phase_us'))"
done
