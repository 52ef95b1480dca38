#!/bin/bash
# One GPU session that regenerates the evidence under profiles/ (run as: gpurun --timeout 3000 -- 'bash tools/gpu_session.sh'):
# whole GPU suite, parity margins, launch lists of the training step and the scaled configuration, ncu --set full of
# every kernel of the step and of the bench.py generation, the bench line.  Summaries: tools/ncu_summary.py.
set -u
mkdir -p gpurun_out
timeout -s KILL 1400 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/r2m_gpu_tests.log 2>&1
tail -4 gpurun_out/r2m_gpu_tests.log | cut -c1-300
timeout -s KILL 900 python tools/parity_margins.py > gpurun_out/parity_margins_r2.txt 2>&1
tail -3 gpurun_out/parity_margins_r2.txt | cut -c1-200
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_train_r2.csv python bench.py --steps 2 --warmup 3 --no-generate --no-cpu-baseline --no-other-configs > gpurun_out/launches_train_r2.log 2>&1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 580 --csv --log-file gpurun_out/launches_scaled_r2.csv python tools/bench_configs.py 4 > gpurun_out/launches_scaled_r2.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_sm100|attn_|layernorm|bias_grad" -f -o gpurun_out/prof_step_r2 python tools/microbench.py --once > gpurun_out/r2m_ncu_step.log 2>&1
tail -2 gpurun_out/r2m_ncu_step.log
# (gpurun brings back at most 64 MiB: the summary is made here and the 60 MB report stays on the box)
cp profiles/ncu_summary.json gpurun_out/ncu_summary_r2.json
python tools/ncu_summary.py kernels gpurun_out/prof_step_r2.ncu-rep gpurun_out/ncu_summary_r2.json "prof_step_r2: tools/microbench.py --once at B 32 T 2048 H 16 d_h 16 (one launch per kernel, ncu --set full --clock-control none)" > /dev/null
rm -f gpurun_out/prof_step_r2.ncu-rep
timeout -s KILL 900 ncu --set full --clock-control none -k regex:decode_mega_kernel -f -o gpurun_out/prof_mega_r2 python tools/profile_decode_mega.py 256 1 1024 > gpurun_out/r2m_ncu_mega.log 2>&1
tail -2 gpurun_out/r2m_ncu_mega.log
python tools/ncu_summary.py kernels gpurun_out/prof_mega_r2.ncu-rep gpurun_out/ncu_summary_r2.json "prof_mega_r2: the bench.py generation exactly (256 sequences, prompt 1, 1,024 events), ncu --set full" > /dev/null
# the decode kernel at the 8-GPU split of the bench (32 sequences per GPU): ncu capture and phase profile
timeout -s KILL 600 ncu --set full --clock-control none -k regex:decode_mega_kernel -f -o gpurun_out/prof_mega_b32_r2 python tools/profile_decode_mega.py 32 1 1024 > gpurun_out/r2m_ncu_mega_b32.log 2>&1
python tools/ncu_summary.py kernels gpurun_out/prof_mega_b32_r2.ncu-rep gpurun_out/ncu_summary_r2.json "prof_mega_b32_r2: generation of 32 sequences (one GPU's share at N = 8), prompt 1, 1,024 events, ncu --set full" > /dev/null
rm -f gpurun_out/prof_mega_b32_r2.ncu-rep gpurun_out/prof_mega_r2.ncu-rep
timeout -s KILL 300 python tools/bench_decode.py 32 1024 8 > gpurun_out/decode_b32_r2.txt 2>&1
timeout -s KILL 300 python tools/bench_decode.py 256 1024 0 > gpurun_out/decode_b256_r2.txt 2>&1
timeout -s KILL 300 python tools/microbench.py > gpurun_out/microbench_r2.txt 2>&1
cp gpurun_out/ncu_summary_r2.json profiles/ncu_summary.json
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err
tail -c 400 gpurun_out/bench_r2_n1.json
