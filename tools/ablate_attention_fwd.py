'''Timing-only ablations of the tcgen05 attention forward (results of the ablated runs are wrong by construction).'''
import ctypes, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from composer_b200 import _lib

def main():
    B, T, H, D = 32, 2048, 16, 16
    E = H * D
    dev = torch.device('cuda:0')
    qkv = torch.randn(B, T, 3 * E, device=dev).to(torch.bfloat16)
    out = torch.empty(B, T, E, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, T, device=dev)
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    names = {0: 'full kernel', 101: 'no max pass', 102: 'no MUFU', 103: 'no max, no MUFU', 104: 'no P store', 108: 'no O~ load', 115: 'none of them'}
    variants = [int(v) for v in sys.argv[1:]] or list(names)
    for impl in variants:
        _lib.call('cb200_set_attention_fwd_impl', impl)
        fn = lambda: _lib.call('cb200_attention_fwd', ptr(qkv), ptr(out), ptr(lse), B, T, H, D, 1 / math.sqrt(D), 0.0, 1, 1, 1, stream)
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): fn()
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) / 5 * 1e3
        print('%-28s %8.1f us  %6.0f cycles / 128x128 tile / SM' % (names.get(impl, str(impl)), us, us * 1.965e3 / (136 * 512 / 148)))
    _lib.call('cb200_set_attention_fwd_impl', 0)

main()
