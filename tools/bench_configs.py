'''
Times the other configurations of BASELINE.json (the contract benchmark, bench.py, is configs[1]):

  configs[0]  forward + loss, default_config.yml hyperparameters, B 4 x T 1024
  configs[3]  long-context training step, T 4096, B 16 per GPU
  configs[4]  scaled Transformer (12 layers, d_model 1024, 16 heads) training step, T 1024, B 16 per GPU

One JSON line per configuration (tokens/s, ms per step, algorithmic TFLOP/s).  Single GPU.
'''

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import VOCAB, load_peaks, step_flops_per_token  # noqa: E402
from composer_b200.models.transformer import Transformer  # noqa: E402


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(steps):
        fn()
    end.record()
    torch.cuda.synchronize()
    return start.elapsed_time(end) * 1e-3 / steps


def run(name, layers, embedding, heads, T, B, train, steps=5):
    model = Transformer(VOCAB, embedding, T, layers, heads, False, 0.0, 0.02, 0.1, 0.1, 1e-5, True, True, seed=0)
    rng = np.random.default_rng(0)
    draw = torch.from_numpy(rng.integers(0, VOCAB, size=(B, T + 1)).astype(np.int32)).cuda()
    x, y = draw[:, :-1].contiguous(), draw[:, 1:].contiguous()
    if train:
        seconds = timed(lambda: model.train_step(x, y), steps)
    else:
        seconds = timed(lambda: model.forward_loss(x, y, training=False), steps)
    flops = step_flops_per_token(layers, embedding, VOCAB, T) / (1 if train else 3)
    peaks, _ = load_peaks()
    tokens_per_s = B * T / seconds
    line = {'config': name, 'layers': layers, 'embedding_size': embedding, 'heads': heads, 'seq_len': T,
            'per_gpu_batch': B, 'phase': 'train step' if train else 'forward + loss', 'tokens_per_s': tokens_per_s,
            'ms_per_step': seconds * 1e3, 'algorithmic_tflops': tokens_per_s * flops / 1e12,
            'frac_of_sustained_bf16_peak': tokens_per_s * flops / 1e12 / peaks['bf16_tflops_sustained'],
            'dropout': 0.1 if train else 0.0}
    print(json.dumps(line), flush=True)
    del model
    torch.cuda.empty_cache()


if __name__ == '__main__':
    # optional arguments: the configurations to run (0, 1, 3, 4; default all) -- e.g. `bench_configs.py 4` under ncu
    only = {int(v) for v in sys.argv[1:]} or {0, 1, 3, 4}
    if 0 in only:
        run('configs[0] forward+loss default', 8, 256, 16, 1024, 4, train=False)
    if 1 in only:
        run('configs[1] train step T2048 (bench.py headline)', 8, 256, 16, 2048, 32, train=True)
    if 3 in only:
        run('configs[3] long context T4096', 8, 256, 16, 4096, 16, train=True)
    if 4 in only:
        run('configs[4] scaled L12 E1024', 12, 1024, 16, 1024, 16, train=True)
