#!/bin/bash
# the three big 8-GPU configurations (BASELINE configs 4, 5 and the Hilbert north-star size)
set -u
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
timeout 600 $TR --master-port 29650 bench.py --gpus 8 --workload hilbert-131072 --steps 3 --warmup 3 --no-e2e 2>gpurun_out/scale_err_h131072.txt | grep '^{' > gpurun_out/scale_hilbert131072_n8.json
timeout 600 $TR --master-port 29651 bench.py --gpus 8 --workload uniform-131072 --steps 1 --warmup 3 --no-e2e 2>gpurun_out/scale_err_u131072.txt | grep '^{' > gpurun_out/scale_uniform131072_n8.json
timeout 600 $TR --master-port 29652 bench.py --gpus 8 --workload uniform-65536 --steps 1 --warmup 3 --no-e2e 2>gpurun_out/scale_err_u65536.txt | grep '^{' > gpurun_out/scale_uniform65536_n8.json
for f in hilbert131072 uniform131072 uniform65536; do python - gpurun_out/scale_${f}_n8.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(d['config']['workload'], 'value', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'], 'rounds', d['rounds'], 'us/round', d['us_per_round'], d['phase_us'], d['clocks'])
PY
done
