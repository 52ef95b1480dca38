'''
Times the individual kernels at the benchmark shapes (config 2 of BASELINE.json
by default: B 32, T 2048, E 256, H 16) with CUDA events and prints achieved
TFLOP/s / GB/s.  Diagnostic tool; bench.py is the contract benchmark.

    python tools/microbench.py [B T [E H]]
'''

import ctypes
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from composer_b200 import _lib  # noqa: E402


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


ONCE = '--once' in sys.argv      # one launch per kernel, no timing: for `ncu --set full` over every kernel of the step
if ONCE:
    sys.argv.remove('--once')


def timeit(fn, iters=10, warmup=3):
    if ONCE:
        fn()
        torch.cuda.synchronize()
        return 1.0
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(iters):
        fn()
    end.record()
    torch.cuda.synchronize()
    return start.elapsed_time(end) / iters * 1e-3


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    E = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    H = int(sys.argv[4]) if len(sys.argv) > 4 else 16
    V = 390
    D, F, M = E // H, 4 * E, B * T
    dev = 'cuda'
    bf = torch.bfloat16
    rows = []

    def gemm(kind, m, n, k, name, extra_bytes=0):
        if kind == 4:
            a = torch.randn(k, m, device=dev).to(bf)
            b = torch.randn(k, n, device=dev).to(bf)
            outf = torch.zeros(m, n, device=dev)
            fn = lambda: _lib.call('cb200_gemm', 4, m, n, k, ptr(a), m, ptr(b), n, None, None, 0, None, 0, None, 0,
                                   ptr(outf), n, 0.0, 0, 0, 0, 0, stream())
            nbytes = (a.numel() + b.numel()) * 2
        else:
            a = torch.randn(m, k, device=dev).to(bf)
            b = torch.randn(n, k, device=dev).to(bf) * 0.05
            bias = torch.randn(n, device=dev)
            out0 = torch.empty(m, n, device=dev, dtype=bf)
            out1 = torch.empty(m, n, device=dev, dtype=bf) if kind == 1 else None
            aux = torch.randn(m, n, device=dev).to(bf) if kind in (2, 3) else None
            rate = 0.1 if kind == 2 else 0.0
            fn = lambda: _lib.call('cb200_gemm', kind, m, n, k, ptr(a), k, ptr(b), k, ptr(bias), ptr(out0), n,
                                   ptr(out1), n, ptr(aux), n, None, 0, rate, 1, 1, 3, 1, stream())
            nbytes = (a.numel() + b.numel() + out0.numel() * (2 if kind == 1 else 1) +
                      (aux.numel() if aux is not None else 0)) * 2
        t = timeit(fn)
        rows.append((name, t * 1e6, 2.0 * m * n * k / t / 1e12, nbytes / t / 1e9))

    gemm(0, M, 3 * E, E, 'c_attn fwd (bias)')
    gemm(2, M, E, E, 'attn c_proj fwd (bias+drop+res)')
    gemm(1, M, F, E, 'c_fc fwd (bias+gelu, 2 outputs)')
    gemm(2, M, E, F, 'mlp c_proj fwd (bias+drop+res)')
    gemm(3, M, F, E, 'mlp c_proj dgrad (*gelu\')')
    gemm(0, M, E, F, 'c_fc dgrad')
    gemm(0, M, E, 3 * E, 'c_attn dgrad')
    gemm(4, E, 3 * E, M, 'c_attn wgrad')
    gemm(4, F, E, M, 'mlp c_proj wgrad')
    gemm(4, E, F, M, 'c_fc wgrad')
    gemm(4, E, E, M, 'attn c_proj wgrad')

    # logits + CE
    h = torch.randn(M, E, device=dev).to(bf)
    wte = (torch.randn(V, E, device=dev) * 0.05).to(bf)
    labels = torch.randint(0, V, (M,), device=dev, dtype=torch.int32)
    dlogits = torch.empty(M, 400, device=dev, dtype=bf)
    loss = torch.zeros(1, device=dev)
    correct = torch.zeros(1, device=dev, dtype=torch.int32)
    t = timeit(lambda: _lib.call('cb200_logits_ce', M, V, E, ptr(h), ptr(wte), ptr(labels), ptr(dlogits), 400,
                                 1.0 / M, ptr(loss), ptr(correct), None, stream()))
    rows.append(('logits + CE (+dlogits)', t * 1e6, 2.0 * M * V * E / t / 1e12, (M * E + M * 400) * 2 / t / 1e9))

    # attention
    scale = 1 / math.sqrt(D)
    qkv = torch.randn(B, T, 3 * E, device=dev).to(bf)
    out = torch.empty(B, T, E, device=dev, dtype=bf)
    lse = torch.empty(B, H, T, device=dev)
    att_flops = 4.0 * B * H * T * (T + 1) / 2 * D
    impl_names = {0: 'tcgen05, P in TMEM', 7: 'tcgen05, two threads per row', 3: 'tcgen05 TS, KT 64, 3 CTA/SM', 4: 'tcgen05 TS, KT 64, 4 CTA/SM', 2: 'tcgen05, P via smem',
                  1: 'mma.sync (round 1)'}
    for impl in (0, 7, 1):
        _lib.call('cb200_set_attention_fwd_impl', impl)
        for rate in (0.0, 0.1):
            try:
                t = timeit(lambda: _lib.call('cb200_attention_fwd', ptr(qkv), ptr(out), ptr(lse), B, T, H, D, scale, rate,
                                             1, 1, 1, stream()), iters=5)
                rows.append(('attention fwd [%s] (dropout %.1f)' % (impl_names[impl], rate), t * 1e6, att_flops / t / 1e12,
                             B * H * T * (T + 1) / 2 / t / 1e12))
            except Exception as error:      # keep the rest of the table
                rows.append(('attention fwd [%s] FAILED: %s' % (impl_names[impl], str(error)[:60]), 0.0, 0.0, 0.0))
    _lib.call('cb200_set_attention_fwd_impl', 0)
    _lib.call('cb200_attention_fwd', ptr(qkv), ptr(out), ptr(lse), B, T, H, D, scale, 0.1, 1, 1, 1, stream())
    dout = torch.randn(B, T, E, device=dev).to(bf)
    delta = torch.empty(B, H, T, device=dev)
    dq_acc = torch.zeros(B, T, E, device=dev)
    dqkv = torch.empty(B, T, 3 * E, device=dev, dtype=bf)
    for rate in (0.0, 0.1):
        t = timeit(lambda: _lib.call('cb200_attention_bwd', ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(delta),
                                     ptr(dq_acc), ptr(dqkv), B, T, H, D, scale, rate, 1, 1, 1, stream()), iters=5)
        rows.append(('attention bwd (dropout %.1f)' % rate, t * 1e6, 2.5 * att_flops / t / 1e12,
                     B * H * T * (T + 1) / 2 / t / 1e12))

    # HBM-bound kernels
    x = torch.randn(M, E, device=dev).to(bf)
    y = torch.empty_like(x)
    gamma = torch.ones(E, device=dev)
    beta = torch.zeros(E, device=dev)
    stats = torch.empty(M, 2, device=dev)
    t = timeit(lambda: _lib.call('cb200_layernorm_fwd', ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(stats), M, E, 1e-5,
                                 stream()))
    rows.append(('layernorm fwd', t * 1e6, 0.0, 2 * M * E * 2 / t / 1e9))
    dg = torch.zeros(E, device=dev)
    db = torch.zeros(E, device=dev)
    t = timeit(lambda: _lib.call('cb200_layernorm_bwd', ptr(x), None, ptr(x), ptr(stats), ptr(gamma), ptr(x), ptr(y),
                                 ptr(dg), ptr(db), M, E, stream()))
    rows.append(('layernorm bwd (+res)', t * 1e6, 0.0, 4 * M * E * 2 / t / 1e9))
    big = torch.randn(M, F, device=dev).to(bf)
    dbias = torch.zeros(F, device=dev)
    t = timeit(lambda: _lib.call('cb200_bias_grad', ptr(big), None, ptr(dbias), M, F, 0.0, 0, 0, 0, 0, stream()))
    rows.append(('bias grad [M, 4E]', t * 1e6, 0.0, M * F * 2 / t / 1e9))
    g_out = torch.empty_like(x)
    dbias2 = torch.zeros(E, device=dev)
    t = timeit(lambda: _lib.call('cb200_bias_grad', ptr(x), ptr(g_out), ptr(dbias2), M, E, 0.1, 1, 1, 3, 1, stream()))
    rows.append(('bias grad + dropout bwd [M, E]', t * 1e6, 0.0, 2 * M * E * 2 / t / 1e9))

    print('shapes: B %d T %d E %d H %d (M = %d tokens)' % (B, T, E, H, M))
    print('%-58s %10s %10s %12s' % ('kernel', 'us', 'TFLOP/s', 'GB/s | Texp/s'))
    for name, us, tf, gb in rows:
        print('%-58s %10.1f %10.1f %12.2f' % (name, us, tf, gb))


if __name__ == '__main__':
    main()
