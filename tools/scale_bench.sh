for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n > gpurun_out/bench_s6_n$n.json 2> gpurun_out/bench_s6_n$n.err
  tail -c 300 gpurun_out/bench_s6_n$n.err | tail -2
done
python bench.py > gpurun_out/bench_s6_n1.json 2> gpurun_out/bench_s6_n1.err
for n in 1 2 4 8; do python - <<PY
import json
d=json.loads(open("gpurun_out/bench_s6_n$n.json").read().strip().splitlines()[-1])
print($n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["generate"]["value"], d["generate"]["us_per_step"])
PY
done
