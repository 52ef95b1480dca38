'''Event timeline of one CTA of the attention backward kernel (cb200_set_attention_trace): prints, per warp, the
SM-clock deltas between consecutive events of the first tiles.  Diagnostic tool.'''
import ctypes, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from composer_b200 import _lib
B, T, H, D = 32, 2048, 16, 16
E = H * D
dev = 'cuda'
ptr = lambda t: ctypes.c_void_p(t.data_ptr())
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
scale = 1 / math.sqrt(D)
rate = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
qkv = torch.randn(B, T, 3 * E, device=dev).to(torch.bfloat16)
out = torch.empty(B, T, E, device=dev, dtype=torch.bfloat16)
dout = torch.randn(B, T, E, device=dev).to(torch.bfloat16)
lse = torch.empty(B, H, T, device=dev)
delta = torch.empty(B, H, T, device=dev)
dq_acc = torch.zeros(B, T, E, device=dev)
dqkv = torch.empty(B, T, 3 * E, device=dev, dtype=torch.bfloat16)
trace = torch.zeros(12 * 512, dtype=torch.int64, device=dev)
_lib.call('cb200_attention_fwd', ptr(qkv), ptr(out), ptr(lse), B, T, H, D, scale, rate, 1, 1, 1, stream)
for _ in range(2):
    _lib.call('cb200_attention_bwd', ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(delta), ptr(dq_acc), ptr(dqkv), B, T, H, D, scale, rate, 1, 1, 1, stream)
_lib.call('cb200_set_attention_trace', ptr(trace))
_lib.call('cb200_attention_bwd', ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(delta), ptr(dq_acc), ptr(dqkv), B, T, H, D, scale, rate, 1, 1, 1, stream)
torch.cuda.synchronize()
_lib.call('cb200_set_attention_trace', None)
t = trace.cpu().numpy().reshape(12, 512)
names = {1: 'start', 2: 'kv landed', 3: 'S/dP(0,0) issued', 10: 'sdp_free(h0) seen', 11: 'S/dP(it,1) issued', 12: 'dq_full(it-1) seen',
         13: 'sdp_free(h1) seen', 14: 'S/dP(it+1,0) issued', 20: 'p_full seen', 21: 'dq_free seen', 22: 'grad MMAs issued',
         30: 's_full h0', 31: 's_full h1', 32: 'pulled h0', 33: 'pulled h1', 34: 'math c0 done', 35: 'dq_full(it-1) seen',
         36: 'stored h0', 37: 'stored h1', 38: 'tile done (dQ drained)', 99: 'end'}
base = min(int(t[w, 1] & 0xFFFFFFFFFF) for w in range(12) if t[w, 0] > 0)
for w in (0, 7, 8, 10):
    n = int(t[w, 0])
    print('--- warp %d: %d events' % (w, n))
    prev = None
    for i in range(1, min(n, 60) + 1):
        ev, clk = int(t[w, i] >> 40), int(t[w, i] & 0xFFFFFFFFFF) - base
        print('  %8d  (+%6d)  %s' % (clk, clk - prev if prev is not None else 0, names.get(ev, ev)))
        prev = clk
end = max(int(t[w, int(t[w, 0])] & 0xFFFFFFFFFF) for w in range(12) if t[w, 0] > 0) - base
print('CTA lifetime %d cycles for 16 tiles = %d per tile' % (end, end // 16))
