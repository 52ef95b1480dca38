#!/bin/bash
# Round-2 GPU session B: attention parity again (dropout keep fraction), ncu --set full of the attention kernels.
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py -q -k "attention or configs" --timeout 600 -p no:cacheprovider > gpurun_out/r2b_attention_tests.log 2>&1
tail -8 gpurun_out/r2b_attention_tests.log
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 4 -c 4 -f -o gpurun_out/prof_attn_r2b python tools/profile_attention.py > gpurun_out/r2b_ncu.log 2>&1
tail -5 gpurun_out/r2b_ncu.log
ls -la gpurun_out/*.ncu-rep
