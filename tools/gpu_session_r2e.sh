#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc -s 1 -c 1 -f -o gpurun_out/prof_attn_r2e python tools/profile_attention.py > gpurun_out/r2e_ncu.log 2>&1
tail -3 gpurun_out/r2e_ncu.log
