#!/bin/bash
# Round 2, call 21 (--gpus 8): cross-GPU round barrier A/B/C in one process group (tools/ab_barrier.py): flat barrier
# with a release/acquire or a sequentially consistent system fence, against the forwarding-flag protocol.
# (Ran at commit d2515d7: the forwarding-flag protocol, its --sweep 17 switch and tools/ab_barrier.py were removed after this
# measurement; check that commit out to repeat it.)
set -u
O=gpurun_out/r2c21; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29557 tools/ab_barrier.py --dims 32768,65536 --solves 8 --laps 2 2> $O/err.txt | grep '^{' > $O/ab_barrier_8gpu.json; echo "rc=$?"
python - $O/ab_barrier_8gpu.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
for r in d["records"]:
    print(r["dim"], r["lap"], r["barrier"], r["us_per_round_median"], r["us_per_round_min"], r["rank0_phase_us"], r["rounds"], r["eigen_val_bits"])
PY
tail -5 $O/err.txt
