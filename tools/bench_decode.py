'''Times cached generation (BASELINE.json configs[2] shape by default) under both decode implementations.

    python tools/bench_decode.py [B [steps [impls...]]]     impls: 1 = per-step graph, 0 = persistent cluster
    kernel (automatic cluster size), 8 / 4 = persistent kernel with that cluster size
'''
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from composer_b200 import _lib  # noqa: E402
from composer_b200.models.transformer import Transformer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
impls = [int(v) for v in sys.argv[3:]] or [1, 0]
E, L, H = 256, 8, 16
model = Transformer(390, E, 1024, L, H, False, 0.0, 0.02, 0.1, 0.1, 1e-5, True, True)
prompt = np.random.default_rng(99).integers(0, 390, size=(B, 1))
weights = 2 * (model.count_params() - 1024 * E + E)
model.generate(prompt[:2], 4)
print('co-resident clusters of the persistent kernel: %d x 8 CTAs, %d x 4 CTAs' % (
    _lib.call('cb200_decode_cluster_capacity', model._engine, 8), _lib.call('cb200_decode_cluster_capacity', model._engine, 4)), flush=True)
for impl in impls:
    _lib.call('cb200_set_decode_impl', 1 if impl == 1 else 0, 0, impl if impl in (4, 8) else 0)
    model.generate(prompt, 16, temperature=1.0, seed=7)
    torch.cuda.synchronize()
    for length in sorted({N // 4, N}):
        start = time.perf_counter()
        out = model.generate(prompt, length, temperature=1.0, seed=7)
        torch.cuda.synchronize()
        seconds = time.perf_counter() - start
        nbytes = weights * length + sum(B * 2 * L * E * 2 * t for t in range(length)) + B * 2 * L * E * 2 * length
        print('impl %d  B %d  steps %4d: %.1f ms, %.1f us/step, %.0f events/s, %.0f GB/s algorithmic, ids %s'
              % (impl, B, length, seconds * 1e3, seconds / length * 1e6, B * length / seconds, nbytes / seconds / 1e9,
                 out[0, :6].tolist()), flush=True)

# phase profile of the persistent kernel (cycles of cluster 0 / CTA 0 / thread 0)
for size in [v for v in impls if v != 1]:
    import ctypes
    names = ['embed', 'ln_1', 'c_attn', 'attention', 'barrier A', 'c_proj', 'barrier B', 'ln_2 + c_fc', 'barrier C',
             'mlp c_proj', 'barrier D', 'ln_f + logits', 'barrier E', 'sample', 'barrier F']
    counters = torch.zeros(64, dtype=torch.int64, device='cuda')
    _lib.call('cb200_set_decode_impl', 0, 0, size if size in (4, 8) else 0)
    _lib.call('cb200_set_decode_profile', ctypes.c_void_p(counters.data_ptr()))
    model.generate(prompt, N, temperature=1.0, seed=7)
    torch.cuda.synchronize()
    _lib.call('cb200_set_decode_profile', None)
    c = counters.cpu().tolist()
    total = sum(c[:16]) or 1
    print('phase profile, cluster size %s, over %d steps (cycles per step; per layer for block phases):' % (size or 'auto', N))
    for i, name in enumerate(names):
        per = c[i] / N / (L if 1 <= i <= 10 else 1)
        print('  %-14s %9.0f cyc  %5.1f%%' % (name, per, 100.0 * c[i] / total))
    print('  (ln_2 alone, up to its __syncthreads: %.0f cyc per layer)' % (c[15] / N / L))
    print('  thread 0 waited for its ring jobs (cyc per layer): c_attn %.0f, attention %.0f, c_proj %.0f, c_fc %.0f, mlp c_proj %.0f; logits %.0f per step'
          % tuple([c[16 + k] / N / L for k in range(5)] + [c[21] / N]))
    print('  inside the linear phases (thread 0, cyc per layer): to run_phase | ring wait | MMAs | ring release | K-split barrier | reduce + epilogue')
    print('  inside the attention phase (thread 0, cyc per layer): q, seed, append %.0f | chunk loop %.0f | merge of the split warps %.0f | output %.0f'
          % tuple(c[54 + i] / N / L for i in range(4)))
    for k, name in enumerate(['c_attn', 'c_proj', 'c_fc', 'mlp c_proj', 'logits']):
        print('    %-12s' % name + ' '.join('%7.0f' % (c[24 + 6 * k + i] / N / (L if k < 4 else 1)) for i in range(6)))
