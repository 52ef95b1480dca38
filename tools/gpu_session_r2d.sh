#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 600 python tools/microbench.py > gpurun_out/r2d_microbench.txt 2>&1
grep -E "attention|shapes" gpurun_out/r2d_microbench.txt
