'''Launches the attention kernels once each at the benchmark shape (for `ncu --set full`).'''
import ctypes, math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from composer_b200 import _lib

B, T, H, D = (int(a) for a in (sys.argv[1:5] + ['32', '2048', '16', '16'][len(sys.argv) - 1:]))
E = H * D
rate = float(sys.argv[5]) if len(sys.argv) > 5 else 0.1
dev = 'cuda'
ptr = lambda t: ctypes.c_void_p(t.data_ptr())
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
scale = 1 / math.sqrt(D)
qkv = torch.randn(B, T, 3 * E, device=dev).to(torch.bfloat16)
out = torch.empty(B, T, E, device=dev, dtype=torch.bfloat16)
dout = torch.randn(B, T, E, device=dev).to(torch.bfloat16)
lse = torch.empty(B, H, T, device=dev)
delta = torch.empty(B, H, T, device=dev)
dq_acc = torch.zeros(B, T, E, device=dev)
dqkv = torch.empty(B, T, 3 * E, device=dev, dtype=torch.bfloat16)
for _ in range(2):
    _lib.call('cb200_attention_fwd', ptr(qkv), ptr(out), ptr(lse), B, T, H, D, scale, rate, 1, 1, 1, stream)
    _lib.call('cb200_attention_bwd', ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(delta), ptr(dq_acc), ptr(dqkv),
              B, T, H, D, scale, rate, 1, 1, 1, stream)
torch.cuda.synchronize()
print('done')
