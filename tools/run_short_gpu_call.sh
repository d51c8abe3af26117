#!/bin/bash
# A GPU call for when only a minute or two of box time is left: no torch import, most informative
# step first, every step writes its own file so a cut-off call still brings something back.
#
#   gpurun --timeout 75 -- 'bash tools/run_short_gpu_call.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $O/short_gpu.txt 2>&1
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > $O/short_smoke.txt 2>&1
echo "rc=$?" >> $O/short_smoke.txt
timeout 120 python -m pytest tests/test_zz_gpu_bitexact.py -q -x -p no:cacheprovider > $O/short_bitexact.txt 2>&1
echo "rc=$?" >> $O/short_bitexact.txt
timeout 120 python -m pytest tests/test_zzz_gpu_bf16_storage.py -q -rxX -p no:cacheprovider > $O/short_bf16.txt 2>&1
echo "rc=$?" >> $O/short_bf16.txt
timeout 120 python -m pytest tests/test_zz_gpu_options_property.py -q -x -p no:cacheprovider > $O/short_options.txt 2>&1
echo "rc=$?" >> $O/short_options.txt
timeout 300 python -m pytest tests -q -m gpu -x -p no:cacheprovider > $O/short_full_suite.txt 2>&1
echo "rc=$?" >> $O/short_full_suite.txt
tail -3 $O/short_*.txt
