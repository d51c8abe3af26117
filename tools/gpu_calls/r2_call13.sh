#!/bin/bash
# Round 2, call 13 (--gpus 2): the tests added after call 11 -- full-size parity against the CPU-computed expectation
# (one GPU and sharded), st_shard_prepare, the wide kernel below the resident limit.
set -u
O=gpurun_out/r2c13; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_bitexact.py tests/test_gpu_sharded.py tests/test_zz_gpu_options_property.py -m gpu -q -rs --durations=8 > $O/pytest.txt 2>&1
tail -25 $O/pytest.txt
