#!/bin/bash
# bench.py at N = $1 GPUs of one box (torchrun for N > 1); writes gpurun_out/bench_r2_n$1.json
n=${1:-8}
if [ "$n" -gt 1 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_r2_n$n.json 2> gpurun_out/bench_r2_n$n.err
else
  python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err
fi
tail -c 300 gpurun_out/bench_r2_n$n.err | tail -2
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r2_n$n.json").read().strip().splitlines()[-1])
print($n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["generate"]["value"], d["generate"]["us_per_step"])
PY
