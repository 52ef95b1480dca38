#!/bin/bash
# Round-2 GPU session C: attention parity for the tile-shape variants + microbench.
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py -q -k "attention" --timeout 600 -p no:cacheprovider > gpurun_out/r2c_attention_tests.log 2>&1
tail -8 gpurun_out/r2c_attention_tests.log
timeout -s KILL 600 python tools/microbench.py > gpurun_out/r2c_microbench.txt 2>&1
grep -E "attention|shapes" gpurun_out/r2c_microbench.txt
