'''Event timeline of CTA 0 of the tcgen05 attention forward (cb200_set_attention_trace).  Diagnostic tool.'''
import ctypes, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from composer_b200 import _lib
B, T, H, D = 32, 2048, 16, 16
E = H * D
dev = 'cuda'
ptr = lambda t: ctypes.c_void_p(t.data_ptr())
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
rate = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
if len(sys.argv) > 2:
    _lib.call('cb200_set_attention_fwd_impl', int(sys.argv[2]))
qkv = torch.randn(B, T, 3 * E, device=dev).to(torch.bfloat16)
out = torch.empty(B, T, E, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, T, device=dev)
trace = torch.zeros(8 * 512, dtype=torch.int64, device=dev)
call = lambda: _lib.call('cb200_attention_fwd', ptr(qkv), ptr(out), ptr(lse), B, T, H, D, 1 / math.sqrt(D), rate, 1, 1, 1, stream)
call(); call()
_lib.call('cb200_set_attention_trace', ptr(trace))
call()
torch.cuda.synchronize()
_lib.call('cb200_set_attention_trace', None)
t = trace.cpu().numpy().reshape(8, 512)
names = {1: 'start', 10: 'S: inputs + buffer ready', 11: 'S issued', 20: 'PV: p_full seen', 21: 'PV issued', 30: 's_full seen',
         31: 'S (+O~ of g-1) in registers', 32: 'max done', 33: 'exps + P stores issued', 34: 'p_full arrived', 35: 'o_full seen'}
base = min(int(t[w, 1] & 0xFFFFFFFFFF) for w in range(8) if t[w, 0] > 0)
for w in (0, 3, 4, 7):
    n = int(t[w, 0])
    print('--- warp %d: %d events' % (w, n))
    prev = None
    for i in range(1, min(n, 70) + 1):
        ev, clk = int(t[w, i] >> 40), int(t[w, i] & 0xFFFFFFFFFF) - base
        print('  %8d  (+%6d)  %s' % (clk, clk - prev if prev is not None else 0, names.get(ev, ev)))
        prev = clk
