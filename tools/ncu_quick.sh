timeout -s KILL 600 ncu --set full --clock-control none -k regex:decode_mega_kernel -f -o gpurun_out/prof_mega_b32 python tools/profile_decode_mega.py 32 1 512 > gpurun_out/ncu_mega_b32.log 2>&1
ncu -i gpurun_out/prof_mega_b32.ncu-rep --page raw --csv > gpurun_out/prof_mega_b32_raw.csv 2>&1
timeout -s KILL 600 ncu --set full --clock-control none -k regex:"layernorm_bwd|bias_grad|layernorm_fwd" -f -o gpurun_out/prof_ew python tools/microbench.py --once > gpurun_out/ncu_ew.log 2>&1
ncu -i gpurun_out/prof_ew.ncu-rep --page raw --csv > gpurun_out/prof_ew_raw.csv 2>&1
ls -la gpurun_out/*.ncu-rep
