'''Runs a short cached generation (for the ncu launch list of the decode step).'''
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from composer_b200.models.transformer import Transformer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
P = int(sys.argv[2]) if len(sys.argv) > 2 else 500
N = int(sys.argv[3]) if len(sys.argv) > 3 else 12
model = Transformer(390, 256, 1024, 8, 16, False, 0.0, 0.02, 0.1, 0.1, 1e-5, True, True)
prompt = np.random.default_rng(0).integers(0, 390, size=(B, P))
out = model.generate(prompt, N, temperature=1.0, seed=3)
torch.cuda.synchronize()
print('done', out.shape)
