import os, sys
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from composer_b200 import _lib
from composer_b200.models.transformer import Transformer
B, P, N = (int(a) for a in (sys.argv[1:4] + ["256", "500", "12"][len(sys.argv) - 1:]))
model = Transformer(390, 256, 1024, 8, 16, False, 0.0, 0.02, 0.1, 0.1, 1e-5, True, True)
prompt = np.random.default_rng(0).integers(0, 390, size=(B, P))
out = model.generate(prompt, N, temperature=1.0, seed=3)
torch.cuda.synchronize()
print('done', out.shape)
