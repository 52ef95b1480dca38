#!/bin/bash
# Round-2 GPU session A: new attention kernels (parity first), microbenchmarks, whole GPU suite, bench line.
set -u
mkdir -p gpurun_out
python -c "import torch; print(torch.cuda.get_device_name(0), torch.version.cuda)" > gpurun_out/r2a_env.txt 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> gpurun_out/r2a_env.txt 2>&1
echo "== attention parity" 
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py -q -k "attention" --timeout 400 -p no:cacheprovider > gpurun_out/r2a_attention_tests.log 2>&1
tail -15 gpurun_out/r2a_attention_tests.log
echo "== microbench"
timeout -s KILL 600 python tools/microbench.py > gpurun_out/r2a_microbench.txt 2>&1
cat gpurun_out/r2a_microbench.txt | tail -32
echo "== gpu suite"
timeout -s KILL 1500 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider -k "not (test_kernel_parity and attention)" > gpurun_out/r2a_gpu_tests.log 2>&1
tail -25 gpurun_out/r2a_gpu_tests.log
echo "== bench"
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 3000 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
