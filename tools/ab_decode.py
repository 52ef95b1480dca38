'''A/B of decode tuning knobs inside one process (same box, same clocks), settings interleaved and repeated.

    python tools/ab_decode.py B LENGTH REPEATS "K1=V1,K2=V2" "K1=V3" ...

Every positional argument after REPEATS is one setting: a comma-separated list of environment variables read by
`decode_mega` at each call (CB200_DECODE_KV_PREFETCH_MB, CB200_DECODE_KV_SPLIT, CB200_DECODE_MAX_CLUSTERS,
CB200_DECODE_L2_HINTS, CB200_DECODE_RING_STAGES, CB200_DECODE_ASYNC_GATHER); "-" is the default configuration.
'''
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from composer_b200.models.transformer import Transformer  # noqa: E402

B, N, R = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
settings = sys.argv[4:] or ['-']
KNOBS = ['CB200_DECODE_KV_PREFETCH_MB', 'CB200_DECODE_KV_SPLIT', 'CB200_DECODE_MAX_CLUSTERS', 'CB200_DECODE_L2_HINTS',
         'CB200_DECODE_RING_STAGES', 'CB200_DECODE_ASYNC_GATHER']
E, L, H = 256, 8, 16
model = Transformer(390, E, 1024, L, H, False, 0.0, 0.02, 0.1, 0.1, 1e-5, True, True)
prompt = np.random.default_rng(99).integers(0, 390, size=(B, 1))


def apply(setting):
    for k in KNOBS:
        os.environ.pop(k, None)
    if setting != '-':
        for kv in setting.split(','):
            k, v = kv.split('=')
            os.environ[k if k.startswith('CB200_') else 'CB200_DECODE_' + k] = v


model.generate(prompt, 32, temperature=1.0, seed=7)
torch.cuda.synchronize()
times = {s: [] for s in settings}
ref = None
for r in range(R):
    for s in settings:
        apply(s)
        start = time.perf_counter()
        out = model.generate(prompt, N, temperature=1.0, seed=7)
        torch.cuda.synchronize()
        times[s].append((time.perf_counter() - start) / N * 1e6)
        out = np.asarray(out.cpu() if hasattr(out, 'cpu') else out)
        if ref is None:
            ref = out
        elif not np.array_equal(ref, out):
            print('  (repeat %d, setting %s: %.4f of the ids differ from the first run)' % (r, s, float((ref != out).mean())))
for s in settings:
    t = sorted(times[s])
    print('B %d  steps %d  %-44s us/step min %.1f  median %.1f  max %.1f   (%.0f events/s at the median)'
          % (B, N, s, t[0], t[len(t) // 2], t[-1], B / t[len(t) // 2] * 1e6), flush=True)
