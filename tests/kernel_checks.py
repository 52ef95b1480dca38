'''
Single-kernel parity checks (GPU).  Each ``check_*`` function runs one kernel
of libcomposer_b200 through the C ABI and compares it with a plain PyTorch
fp32 evaluation of the same operation on the same inputs; it returns a dict
with the error statistics and raises ``AssertionError`` on a mismatch.

Used by ``tests/test_kernels_gpu.py`` (pytest, ``-m gpu``) and by
``tools/gpu_probe.py`` (which runs every check without stopping and writes a
report under ``gpurun_out/``).
'''

import ctypes
import math

import torch

from composer_b200 import _lib

DEV = 'cuda'


def ptr(tensor):
    return ctypes.c_void_p(tensor.data_ptr()) if tensor is not None else None


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _stats(name, got, ref, tol, scale=None):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    denom = float(ref.abs().max()) if scale is None else scale
    denom = max(denom, 1e-30)
    result = {'name': name, 'max_abs_err': float(err.max()), 'ref_max': float(ref.abs().max()),
              'rel': float(err.max()) / denom, 'tol': tol, 'nan': bool(torch.isnan(got).any())}
    result['ok'] = (not result['nan']) and result['rel'] <= tol
    if not result['ok'] and got.dim() == 2:
        # where is it wrong?  error per 64-column slab / 32-row band helps to spot layout bugs
        rows, cols = err.shape
        slabs = [float(err[:, c:c + 64].max()) for c in range(0, cols, 64)][:16]
        bands = [float(err[r:r + 32].max()) for r in range(0, rows, 32)][:16]
        result['err_by_col_slab64'] = ['%.3g' % v for v in slabs]
        result['err_by_row_band32'] = ['%.3g' % v for v in bands]
        idx = int(err.argmax())
        result['worst'] = (idx // cols, idx % cols, float(got.flatten()[idx]), float(ref.flatten()[idx]))
    return result


def _finish(results):
    if isinstance(results, dict):
        results = [results]
    bad = [r for r in results if not r['ok']]
    assert not bad, 'kernel mismatch: %r' % bad
    return results


def _randn(*shape, scale=1.0, dtype=torch.bfloat16, seed=0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(DEV)


# ---------------------------------------------------------------------------
# GEMM family
# ---------------------------------------------------------------------------

def check_gemm_bias(M=256, N=768, K=256, b_mn=False, seed=1):
    a = _randn(M, K, seed=seed)
    w = _randn(K, N, scale=0.05, seed=seed + 1)          # [in, out]
    bias = _randn(N, dtype=torch.float32, seed=seed + 2)
    out = torch.full((M, N), float('nan'), dtype=torch.bfloat16, device=DEV)
    if b_mn:
        b_mat, ldb, kind = w.contiguous(), N, 6           # stored [K, N]
    else:
        b_mat, ldb, kind = w.t().contiguous(), K, 0       # stored [N, K]
    _lib.call('cb200_gemm', kind, M, N, K, ptr(a), K, ptr(b_mat), ldb, ptr(bias), ptr(out), N, None, 0, None, 0,
              None, 0, 0.0, 0, 0, 0, 0, stream())
    torch.cuda.synchronize()
    ref = a.float() @ w.float() + bias
    return _finish(_stats('gemm_bias M%d N%d K%d b_mn=%s' % (M, N, K, b_mn), out, ref, 1e-2))


def check_gemm_gelu(M=256, N=1024, K=256, seed=3):
    a = _randn(M, K, seed=seed)
    w = _randn(K, N, scale=0.08, seed=seed + 1)
    bias = _randn(N, dtype=torch.float32, seed=seed + 2)
    out0 = torch.full((M, N), float('nan'), dtype=torch.bfloat16, device=DEV)
    out1 = torch.full((M, N), float('nan'), dtype=torch.bfloat16, device=DEV)
    _lib.call('cb200_gemm', 1, M, N, K, ptr(a), K, ptr(w.t().contiguous()), K, ptr(bias), ptr(out0), N, ptr(out1), N,
              None, 0, None, 0, 0.0, 0, 0, 0, 0, stream())
    torch.cuda.synchronize()
    pre = a.float() @ w.float() + bias
    gel = 0.5 * pre * (1 + torch.tanh(math.sqrt(2 / math.pi) * (pre + 0.044715 * pre ** 3)))
    return _finish([_stats('gemm_gelu pre', out0, pre, 1e-2), _stats('gemm_gelu act', out1, gel, 1e-2)])


def rowmajor_mask(rows, cols, rate, seed, step, site, layer):
    mask = torch.empty((rows, cols), dtype=torch.uint8, device=DEV)
    _lib.call('cb200_rowmajor_dropout_mask', ptr(mask), rows, cols, rate, seed, step, site, layer, stream())
    return mask


def check_gemm_drop_res(M=256, N=256, K=1024, rate=0.0, seed=5):
    a = _randn(M, K, seed=seed)
    w = _randn(K, N, scale=0.05, seed=seed + 1)
    bias = _randn(N, dtype=torch.float32, seed=seed + 2)
    res = _randn(M, N, seed=seed + 3)
    out = torch.full((M, N), float('nan'), dtype=torch.bfloat16, device=DEV)
    _lib.call('cb200_gemm', 2, M, N, K, ptr(a), K, ptr(w.t().contiguous()), K, ptr(bias), ptr(out), N, None, 0,
              ptr(res), N, None, 0, rate, 77, 3, 4, 2, stream())
    torch.cuda.synchronize()
    y = a.float() @ w.float() + bias
    results = []
    if rate > 0:
        mask = rowmajor_mask(M, N, rate, 77, 3, 4, 2).float()
        torch.cuda.synchronize()
        keep = float(mask.mean())
        results.append({'name': 'dropout keep fraction', 'ok': abs(keep - (1 - rate)) < 0.02, 'keep': keep,
                        'rel': abs(keep - (1 - rate)), 'tol': 0.02, 'nan': False})
        y = y * mask / (1 - rate)
    results.append(_stats('gemm_drop_res rate=%g' % rate, out, res.float() + y, 1e-2))
    return _finish(results)


def check_gemm_dgelu(M=256, N=1024, K=256, seed=9):
    a = _randn(M, K, seed=seed)
    w = _randn(N, K, scale=0.05, seed=seed + 1)           # stored [N, K] directly
    u = _randn(M, N, seed=seed + 2)
    out = torch.full((M, N), float('nan'), dtype=torch.bfloat16, device=DEV)
    _lib.call('cb200_gemm', 3, M, N, K, ptr(a), K, ptr(w), K, None, ptr(out), N, None, 0, ptr(u), N, None, 0,
              0.0, 0, 0, 0, 0, stream())
    torch.cuda.synchronize()
    uf = u.float().requires_grad_(True)
    gel = 0.5 * uf * (1 + torch.tanh(math.sqrt(2 / math.pi) * (uf + 0.044715 * uf ** 3)))
    gel.sum().backward()
    ref = (a.float() @ w.float().t()) * uf.grad
    return _finish(_stats('gemm_dgelu', out, ref, 1e-2))


def check_gemm_wgrad(tokens=512, M=256, N=768, seed=11, ldm_pad=0):
    '''outf[M, N] += A^T B with A [tokens, M], B [tokens, N] (both row-major, token-major).'''

    a = _randn(tokens, M + ldm_pad, seed=seed)
    b = _randn(tokens, N, seed=seed + 1)
    init = _randn(M, N, dtype=torch.float32, seed=seed + 2)
    outf = init.clone()
    _lib.call('cb200_gemm', 4, M, N, tokens, ptr(a), M + ldm_pad, ptr(b), N, None, None, 0, None, 0, None, 0,
              ptr(outf), N, 0.0, 0, 0, 0, 0, stream())
    torch.cuda.synchronize()
    ref = init + a[:, :M].float().t() @ b.float()
    return _finish(_stats('gemm_wgrad tokens%d M%d N%d' % (tokens, M, N), outf, ref, 2e-5))      # fp32 accumulation of the same bf16 operands: measured <= 7e-7


def check_logits_ce(M=300, V=390, E=256, seed=13):
    h = _randn(M, E, seed=seed)
    wte = _randn(V, E, scale=0.1, seed=seed + 1)
    g = torch.Generator().manual_seed(seed + 2)
    labels = torch.randint(0, V, (M,), generator=g, dtype=torch.int32).to(DEV)
    vpad = (V + 15) // 16 * 16
    dlogits = torch.full((M, vpad), float('nan'), dtype=torch.bfloat16, device=DEV)
    logits = torch.full((M, V), float('nan'), dtype=torch.float32, device=DEV)
    loss = torch.zeros(1, dtype=torch.float32, device=DEV)
    correct = torch.zeros(1, dtype=torch.int32, device=DEV)
    scale = 1.0 / M
    _lib.call('cb200_logits_ce', M, V, E, ptr(h), ptr(wte), ptr(labels), ptr(dlogits), vpad, scale, ptr(loss),
              ptr(correct), ptr(logits), stream())
    torch.cuda.synchronize()
    z = (h.float() @ wte.float().t()).requires_grad_(True)
    ref_loss = torch.nn.functional.cross_entropy(z, labels.long(), reduction='sum')
    (ref_loss * scale).backward()
    ref_correct = int((z.argmax(dim=-1) == labels.long()).sum())
    results = [_stats('ce logits', logits, z.detach(), 2e-5),      # fp32 out of TMEM: measured 6e-7
               _stats('ce dlogits', dlogits[:, :V], z.grad, 1e-2),
               _stats('ce dlogits pad', dlogits[:, V:], torch.zeros_like(dlogits[:, V:]), 0.0, scale=1.0),
               _stats('ce loss', loss, ref_loss.detach().reshape(1), 5e-6)]
    hits = int(correct.item())
    results.append({'name': 'ce correct', 'ok': abs(hits - ref_correct) <= 1, 'got': hits, 'ref': ref_correct,
                    'rel': 0.0, 'tol': 0.0, 'nan': False})
    return _finish(results)


# ---------------------------------------------------------------------------
# HBM-bound kernels
# ---------------------------------------------------------------------------

def check_embed(B=3, T=40, E=256, V=390, rate=0.0):
    g = torch.Generator().manual_seed(21)
    ids = torch.randint(0, V, (B, T), generator=g, dtype=torch.int32).to(DEV)
    wte = _randn(V, E, dtype=torch.float32, seed=22)
    wpe = _randn(64, E, dtype=torch.float32, seed=23)
    out = torch.empty((B * T, E), dtype=torch.bfloat16, device=DEV)
    _lib.call('cb200_embed_fwd', ptr(ids), ptr(wte), ptr(wpe), ptr(out), B, T, E, 5, V, rate, 9, 2, stream())
    ref = wte[ids.long()] + wpe[5:5 + T][None]
    mask = None
    if rate > 0:
        mask = rowmajor_mask(B * T, E, rate, 9, 2, 1, 0).float().reshape(B, T, E)
        ref = ref * mask / (1 - rate)
    torch.cuda.synchronize()
    results = [_stats('embed_fwd rate=%g' % rate, out.reshape(B, T, E), ref, 5e-3)]
    # backward
    dh = _randn(B * T, E, seed=24)
    dwte = torch.zeros((V, E), dtype=torch.float32, device=DEV)
    dwpe = torch.zeros((64, E), dtype=torch.float32, device=DEV)
    _lib.call('cb200_embed_bwd', ptr(ids), ptr(dh), ptr(dwte), ptr(dwpe), B, T, E, 5, V, rate, 9, 2, stream())
    torch.cuda.synchronize()
    gsrc = dh.float().reshape(B, T, E)
    if mask is not None:
        gsrc = gsrc * mask / (1 - rate)
    ref_wte = torch.zeros_like(dwte).index_add_(0, ids.long().reshape(-1), gsrc.reshape(-1, E))
    ref_wpe = torch.zeros_like(dwpe)
    ref_wpe[5:5 + T] = gsrc.sum(dim=0)
    results.append(_stats('embed_bwd dwte', dwte, ref_wte, 1e-4))
    results.append(_stats('embed_bwd dwpe', dwpe, ref_wpe, 1e-4))
    return _finish(results)


def check_layernorm(rows=301, E=256):
    x = _randn(rows, E, scale=2.0, seed=31) + 0.5
    gamma = _randn(E, dtype=torch.float32, seed=32) * 0.2 + 1.0
    beta = _randn(E, dtype=torch.float32, seed=33) * 0.1
    y = torch.empty_like(x)
    stats = torch.empty((rows, 2), dtype=torch.float32, device=DEV)
    _lib.call('cb200_layernorm_fwd', ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(stats), rows, E, 1e-5, stream())
    xf = x.float().requires_grad_(True)
    gf = gamma.clone().requires_grad_(True)
    bf = beta.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xf, (E,), gf, bf, 1e-5)
    torch.cuda.synchronize()
    results = [_stats('layernorm_fwd', y, ref.detach(), 1e-2)]
    dy_a = _randn(rows, E, seed=34)
    dy_b = _randn(rows, E, seed=35)
    dres = _randn(rows, E, seed=36)
    dx = torch.empty_like(x)
    dgamma = torch.zeros(E, dtype=torch.float32, device=DEV)
    dbeta = torch.zeros(E, dtype=torch.float32, device=DEV)
    _lib.call('cb200_layernorm_bwd', ptr(dy_a), ptr(dy_b), ptr(x), ptr(stats), ptr(gamma), ptr(dres), ptr(dx),
              ptr(dgamma), ptr(dbeta), rows, E, stream())
    torch.cuda.synchronize()
    ref.backward(dy_a.float() + dy_b.float())
    results.append(_stats('layernorm_bwd dx', dx, xf.grad + dres.float(), 1e-2))
    results.append(_stats('layernorm_bwd dgamma', dgamma, gf.grad, 1e-5))     # fp32 column sums: measured 2e-7
    results.append(_stats('layernorm_bwd dbeta', dbeta, bf.grad, 1e-5))
    return _finish(results)


def check_bias_grad(rows=1000, N=768, rate=0.1):
    dy = _randn(rows, N, seed=41)
    g_out = torch.empty_like(dy)
    dbias = torch.zeros(N, dtype=torch.float32, device=DEV)
    _lib.call('cb200_bias_grad', ptr(dy), ptr(g_out), ptr(dbias), rows, N, rate, 5, 1, 3, 2, stream())
    torch.cuda.synchronize()
    ref = dy.float()
    if rate > 0:
        ref = ref * rowmajor_mask(rows, N, rate, 5, 1, 3, 2).float() / (1 - rate)
        torch.cuda.synchronize()
    results = [_stats('bias_grad dbias', dbias, ref.sum(dim=0), 1e-5)]       # measured <= 3e-7
    if rate > 0:
        results.append(_stats('bias_grad g_out', g_out, ref, 1e-2))
    return _finish(results)


def check_adam(n=100000):
    p = _randn(n, dtype=torch.float32, seed=51)
    g = _randn(n, dtype=torch.float32, seed=52) * 0.01
    m = _randn(n, dtype=torch.float32, seed=53) * 0.01
    v = (_randn(n, dtype=torch.float32, seed=54) * 0.01) ** 2
    shadow = torch.empty(n, dtype=torch.bfloat16, device=DEV)
    p0, m0, v0 = p.clone(), m.clone(), v.clone()
    lr_t, b1, b2, eps, gs = 1e-3, 0.9, 0.999, 1e-7, 0.5
    _lib.call('cb200_adam', ptr(p), ptr(g), ptr(m), ptr(v), ptr(shadow), n, lr_t, b1, b2, eps, gs, stream())
    torch.cuda.synchronize()
    gg = g * gs
    mr = b1 * m0 + (1 - b1) * gg
    vr = b2 * v0 + (1 - b2) * gg * gg
    pr = p0 - lr_t * mr / (vr.sqrt() + eps)
    return _finish([_stats('adam p', p, pr, 1e-5), _stats('adam m', m, mr, 1e-5), _stats('adam v', v, vr, 1e-5),
                    _stats('adam shadow', shadow, pr, 1e-2)])


# ---------------------------------------------------------------------------
# Attention
# ---------------------------------------------------------------------------

def attention_mask(B, T, H, rate, seed, step, layer):
    mask = torch.empty((B, H, T, T), dtype=torch.uint8, device=DEV)
    _lib.call('cb200_attention_dropout_mask', ptr(mask), B, T, H, rate, seed, step, layer, stream())
    return mask


def _attention_reference(qkv, B, T, H, D, scale, keep=None, rate=0.0):
    E = H * D
    q, k, v = qkv.float().reshape(B, T, 3, H, D).permute(2, 0, 3, 1, 4)
    w = (q @ k.transpose(-1, -2)) * scale
    causal = torch.tril(torch.ones(T, T, device=qkv.device))
    w = w * causal - 1e4 * (1 - causal)               # transformer.py:351-354
    p = torch.softmax(w, dim=-1)
    if keep is not None:
        p = p * keep.float() / (1 - rate)
    out = (p @ v).permute(0, 2, 1, 3).reshape(B, T, E)
    return out


def check_attention(B=2, T=200, H=16, D=16, rate=0.0, backward=True, fwd_impl=0, sharp=1.0):
    '''fwd_impl: 0 = tcgen05 with P in TMEM (default), 1 = round-1 mma.sync kernel, 2 = tcgen05 with P through smem,
    3 / 4 = tile-shape variants of 0 (d_h 16), 7 = 0 with two threads per score row.  ``sharp`` scales the scores
    (their standard deviation, in nats).'''
    _lib.call('cb200_set_attention_fwd_impl', fwd_impl)
    try:
        return _check_attention(B, T, H, D, rate, backward, sharp)
    finally:
        _lib.call('cb200_set_attention_fwd_impl', 0)


def _check_attention(B, T, H, D, rate, backward, sharp=1.0):
    E = H * D
    scale = 1.0 / math.sqrt(D)
    qkv = _randn(B, T, 3 * E, scale=1.0, seed=61)
    if sharp != 1.0:
        qkv[..., :2 * E] = (qkv[..., :2 * E].float() * math.sqrt(sharp)).to(torch.bfloat16)
    out = torch.full((B, T, E), float('nan'), dtype=torch.bfloat16, device=DEV)
    lse = torch.empty((B, H, T), dtype=torch.float32, device=DEV)
    _lib.call('cb200_attention_fwd', ptr(qkv), ptr(out), ptr(lse), B, T, H, D, scale, rate, 123, 7, 3, stream())
    torch.cuda.synchronize()
    keep = attention_mask(B, T, H, rate, 123, 7, 3) if rate > 0 else None
    qf = qkv.float().requires_grad_(True)
    ref = _attention_reference(qf, B, T, H, D, scale, keep, rate)
    results = [_stats('attention_fwd B%d T%d H%d D%d rate=%g' % (B, T, H, D, rate), out.reshape(B * T, E),
                      ref.detach().reshape(B * T, E), 1.5e-2)]
    if keep is not None:
        lower = torch.tril(torch.ones(T, T, device=DEV)).bool()
        frac = float(keep[..., lower].float().mean())
        results.append({'name': 'attention keep fraction', 'ok': abs(frac - (1 - rate)) < 0.02, 'keep': frac,
                        'rel': abs(frac - (1 - rate)), 'tol': 0.02, 'nan': False})
    if backward:
        dout = _randn(B, T, E, seed=62)
        ref.backward(dout.float())
        delta = torch.empty((B, H, T), dtype=torch.float32, device=DEV)
        dq_acc = torch.zeros((B, T, E), dtype=torch.float32, device=DEV)
        dqkv = torch.full((B, T, 3 * E), float('nan'), dtype=torch.bfloat16, device=DEV)
        _lib.call('cb200_attention_bwd', ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(delta), ptr(dq_acc), ptr(dqkv),
                  B, T, H, D, scale, rate, 123, 7, 3, stream())
        torch.cuda.synchronize()
        g = qf.grad.reshape(B * T, 3 * E)
        d = dqkv.reshape(B * T, 3 * E)
        for i, nm in enumerate(('dq', 'dk', 'dv')):
            results.append(_stats('attention_bwd %s' % nm, d[:, i * E:(i + 1) * E], g[:, i * E:(i + 1) * E], 3e-2))
        results.append(_stats('attention_bwd dq_acc rezeroed', dq_acc.reshape(B * T, E),
                              torch.zeros(B * T, E, device=DEV), 0.0, scale=1.0))
    return _finish(results)


def check_decode_attention(B=3, H=16, D=16, t_max=96, pos=70):
    E = H * D
    scale = 1.0 / math.sqrt(D)
    kc = _randn(B, H, t_max, D, seed=71)
    vc = _randn(B, H, t_max, D, seed=72)
    qkv = _randn(B, 3 * E, seed=73)
    out = torch.empty((B, E), dtype=torch.bfloat16, device=DEV)
    pos_t = torch.tensor([pos, 0], dtype=torch.int32, device=DEV)
    kc2, vc2 = kc.clone(), vc.clone()
    _lib.call('cb200_decode_attention', ptr(qkv), ptr(kc2), ptr(vc2), ptr(out), ptr(pos_t), B, H, D, t_max, scale,
              stream())
    torch.cuda.synchronize()
    q = qkv[:, :E].float().reshape(B, H, D)
    kn = qkv[:, E:2 * E].reshape(B, H, D)
    vn = qkv[:, 2 * E:].reshape(B, H, D)
    kr, vr = kc.clone(), vc.clone()
    kr[:, :, pos] = kn
    vr[:, :, pos] = vn
    s = torch.einsum('bhd,bhtd->bht', q, kr[:, :, :pos + 1].float()) * scale
    p = torch.softmax(s, dim=-1)
    ref = torch.einsum('bht,bhtd->bhd', p, vr[:, :, :pos + 1].float()).reshape(B, E)
    return _finish([_stats('decode_attention', out, ref, 1e-2),
                    _stats('decode k append', kc2.reshape(-1, D), kr.reshape(-1, D), 0.0, scale=1.0),
                    _stats('decode v append', vc2.reshape(-1, D), vr.reshape(-1, D), 0.0, scale=1.0)])


def check_decode_linear(B=37, N=768, K=256, epilogue=0, seed=81):
    x = _randn(B, K, seed=seed)
    w = _randn(N, K, scale=0.05, seed=seed + 1)          # stored [N, K]
    bias = _randn(N, dtype=torch.float32, seed=seed + 2)
    res = _randn(B, N, seed=seed + 3)
    y = torch.full((B, N), float('nan'), dtype=torch.bfloat16, device=DEV)
    _lib.call('cb200_decode_linear', epilogue, ptr(x), K, ptr(w), ptr(bias), ptr(res), N, ptr(y), N, B, N, K, stream())
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t() + bias
    if epilogue == 1:
        ref = 0.5 * ref * (1 + torch.tanh(math.sqrt(2 / math.pi) * (ref + 0.044715 * ref ** 3)))
    if epilogue == 2:
        ref = ref + res.float()
    return _finish(_stats('decode_linear B%d N%d K%d epi%d' % (B, N, K, epilogue), y, ref, 1e-2))


GROUPS = {
    'gemm_basic': [lambda: check_gemm_bias(128, 256, 64), lambda: check_gemm_bias(256, 768, 256),
                   lambda: check_gemm_bias(300, 768, 256), lambda: check_gemm_bias(4096, 256, 1024),
                   lambda: check_gemm_bias(20000, 768, 256)],
    'gemm_bmn': [lambda: check_gemm_bias(128, 256, 64, b_mn=True), lambda: check_gemm_bias(300, 768, 256, b_mn=True)],
    'gemm_wgrad': [lambda: check_gemm_wgrad(64, 128, 256), lambda: check_gemm_wgrad(512, 256, 768),
                   lambda: check_gemm_wgrad(8192, 1024, 256), lambda: check_gemm_wgrad(1000, 390, 256, ldm_pad=10)],
    'gemm_epilogues': [check_gemm_gelu, lambda: check_gemm_drop_res(rate=0.0), lambda: check_gemm_drop_res(rate=0.1),
                       check_gemm_dgelu],
    'logits_ce': [check_logits_ce, lambda: check_logits_ce(M=1024, V=390), lambda: check_logits_ce(M=130, V=500)],
    'elementwise': [check_embed, lambda: check_embed(rate=0.1), check_layernorm,
                    lambda: check_layernorm(rows=77, E=1024), check_bias_grad, lambda: check_bias_grad(rate=0.0), lambda: check_bias_grad(500, 3072, 0.0),
                    lambda: check_bias_grad(300, 4096, 0.1), lambda: check_bias_grad(64, 256, 0.1), lambda: check_bias_grad(5, 1536, 0.0),
                    check_adam],
    'attention_fwd': [lambda: check_attention(1, 64, 2, 16, backward=False),
                      lambda: check_attention(2, 200, 16, 16, backward=False),
                      lambda: check_attention(2, 256, 4, 64, backward=False),
                      lambda: check_attention(2, 200, 16, 16, rate=0.1, backward=False),
                      lambda: check_attention(1, 130, 4, 32, rate=0.1, backward=False),
                      # more work items than resident CTAs (persistent loop, Q / K / V rings wrap, both S buffers)
                      lambda: check_attention(6, 1024, 16, 16, rate=0.1, backward=False),
                      lambda: check_attention(2, 640, 16, 64, rate=0.1, backward=False),
                      # sharp attention (score standard deviation 24 / 12 nats): running maxima far above the first tile's
                      lambda: check_attention(2, 1024, 4, 16, rate=0.1, backward=False, sharp=24.0),
                      lambda: check_attention(1, 512, 2, 64, backward=False, sharp=12.0),
                      # two threads per score row (8 softmax warps per CTA)
                      lambda: check_attention(2, 200, 16, 16, rate=0.1, backward=False, fwd_impl=7),
                      lambda: check_attention(6, 1024, 16, 16, rate=0.1, backward=False, fwd_impl=7),
                      lambda: check_attention(2, 1024, 4, 16, backward=False, sharp=24.0, fwd_impl=7),
                      # the other two forward implementations: P staged through shared memory, round-1 mma.sync kernel
                      lambda: check_attention(2, 200, 16, 16, rate=0.1, backward=False, fwd_impl=2),
                      lambda: check_attention(6, 1024, 16, 16, backward=False, fwd_impl=2),
                      lambda: check_attention(2, 384, 4, 64, rate=0.1, backward=False, fwd_impl=2),
                      lambda: check_attention(1, 192, 4, 32, rate=0.1, backward=False, fwd_impl=2),
                      # tile-shape variants of the default (64-key tiles, 3 / 4 CTAs per SM)
                      lambda: check_attention(2, 200, 16, 16, rate=0.1, backward=False, fwd_impl=3),
                      lambda: check_attention(6, 1024, 16, 16, rate=0.1, backward=False, fwd_impl=3),
                      lambda: check_attention(2, 200, 16, 16, rate=0.1, backward=False, fwd_impl=4),
                      lambda: check_attention(6, 1024, 16, 16, backward=False, fwd_impl=4),
                      lambda: check_attention(2, 200, 16, 16, rate=0.1, backward=False, fwd_impl=1),
                      lambda: check_attention(2, 256, 4, 64, rate=0.1, backward=False, fwd_impl=1),
                      lambda: check_attention(1, 192, 4, 32, rate=0.1, backward=False, fwd_impl=1)],
    'attention_bwd': [lambda: check_attention(1, 64, 2, 16), lambda: check_attention(2, 200, 16, 16),
                      lambda: check_attention(1, 256, 4, 64), lambda: check_attention(1, 192, 4, 32),
                      lambda: check_attention(2, 200, 16, 16, rate=0.1),
                      lambda: check_attention(1, 4096, 2, 16, rate=0.1), lambda: check_attention(1, 1000, 2, 64, rate=0.1),
                      lambda: check_attention(3, 130, 4, 32, rate=0.1), lambda: check_attention(2, 16, 4, 16),
                      lambda: check_attention(1, 2048, 4, 16, rate=0.1)],
    'decode_linear': [check_decode_linear, lambda: check_decode_linear(256, 1024, 256, 1),
                      lambda: check_decode_linear(130, 256, 1024, 2), lambda: check_decode_linear(1, 256, 256, 2),
                      lambda: check_decode_linear(32, 3072, 1024, 0)],
    'decode': [check_decode_attention, lambda: check_decode_attention(2, 4, 64, 300, 299),
               lambda: check_decode_attention(2, 16, 16, 64, 0)],
}


# ---------------------------------------------------------------------------
# Whole model against the CPU oracle (oracle/transformer_oracle.py)
# ---------------------------------------------------------------------------

def _small_model(layers=2, embedding=256, heads=16, window=128, vocab=390, dropout=0.0, seed=3, layer_norm=True):
    from composer_b200.models.transformer import Transformer
    from oracle import transformer_oracle as oracle

    cfg = oracle.OracleConfig(vocab_size=vocab, embedding_size=embedding, window_size=window,
                              decoder_layers_count=layers, attention_head_count=heads,
                              attention_dropout_rate=dropout, residual_dropout_rate=dropout,
                              use_layer_normalization=layer_norm)
    weights = oracle.init_parameters(cfg, seed=seed)
    # make biases / LayerNorm parameters non-trivial so that their gradients and uses are exercised
    rng = __import__('numpy').random.default_rng(seed + 1)
    for name in weights:
        if name.endswith('/bias') or name.endswith('/beta'):
            weights[name] = (0.02 * rng.standard_normal(weights[name].shape)).astype('float32')
        elif name.endswith('/gamma'):
            weights[name] = (1.0 + 0.05 * rng.standard_normal(weights[name].shape)).astype('float32')
    model = Transformer(vocab, embedding, window, layers, heads, False, 0.0, 0.02, dropout, dropout, 1e-5, True, layer_norm)
    model.set_weights(weights)
    return model, cfg, weights


def check_engine_forward_backward(B=2, T=100, layers=2, embedding=256, heads=16, tol=2e-2, layer_norm=True,
                                  oracle_dtype=None):
    import numpy as np
    from oracle import transformer_oracle as oracle

    model, cfg, weights = _small_model(layers, embedding, heads, window=max(T, 128), layer_norm=layer_norm)
    rng = np.random.default_rng(11)
    draw = rng.integers(0, cfg.vocab_size, size=(B, T + 1))
    x, y = draw[:, :-1], draw[:, 1:]
    loss_sum, correct, logits = model.forward_loss(x, y, training=True, return_logits=True)
    model.backward()
    torch.cuda.synchronize()
    ref_loss, ref_acc, ref_logits, ref_grads = oracle.loss_and_gradients(weights, x, y, cfg,
                                                                          dtype=oracle_dtype or torch.float64)
    results = []
    got_logits = logits.cpu().double().numpy().reshape(B * T, -1)
    results.append(_stats('engine logits', torch.from_numpy(got_logits), torch.from_numpy(ref_logits.reshape(B * T, -1)), tol))
    got_loss = float(loss_sum) / (B * T)
    # (the mean over all positions averages the bf16 noise of the logits away: measured <= 4e-5)
    results.append({'name': 'engine loss', 'got': got_loss, 'ref': ref_loss, 'rel': abs(got_loss - ref_loss) / abs(ref_loss),
                    'tol': 5e-4, 'nan': got_loss != got_loss, 'ok': abs(got_loss - ref_loss) <= 5e-4 * abs(ref_loss)})
    got_acc = float(correct) / (B * T)
    results.append({'name': 'engine accuracy', 'got': got_acc, 'ref': ref_acc, 'rel': abs(got_acc - ref_acc), 'tol': 0.02,
                    'nan': False, 'ok': abs(got_acc - ref_acc) <= 0.02})
    grads = model.get_gradients()
    for name, ref in ref_grads.items():
        g = torch.from_numpy(np.asarray(grads[name], dtype=np.float64).reshape(ref.shape))
        r = torch.from_numpy(np.asarray(ref))
        denom = float(r.norm()) + 1e-12
        rel = float((g - r).norm()) / denom
        results.append({'name': 'grad ' + name, 'rel': rel, 'tol': 5e-2, 'ref_norm': denom, 'nan': bool(torch.isnan(g).any()),
                        'ok': rel <= 5e-2 and not bool(torch.isnan(g).any())})
    return _finish(results)


def check_engine_adam_step(B=2, T=64):
    import numpy as np
    from oracle import transformer_oracle as oracle

    model, cfg, weights = _small_model(2, 256, 16)
    rng = np.random.default_rng(5)
    draw = rng.integers(0, cfg.vocab_size, size=(B, T + 1))
    x, y = draw[:, :-1], draw[:, 1:]
    state = oracle.AdamState(weights, learning_rate=1e-3)
    params = {k: v.copy() for k, v in weights.items()}
    losses_ref, losses_got = [], []
    for step in range(3):
        ref_loss, _, _, ref_grads = oracle.loss_and_gradients(params, x, y, cfg, dtype=torch.float64)
        params = state.apply(params, ref_grads)
        losses_ref.append(ref_loss)
        loss_sum, _ = model.train_step(x, y, 1e-3)
        losses_got.append(float(loss_sum) / (B * T))
    torch.cuda.synchronize()
    results = []
    for i, (a, b) in enumerate(zip(losses_got, losses_ref)):
        results.append({'name': 'adam trajectory loss[%d]' % i, 'got': a, 'ref': b, 'rel': abs(a - b) / abs(b), 'tol': 5e-4,
                        'nan': a != a, 'ok': abs(a - b) <= 5e-4 * abs(b)})      # measured <= 2.2e-5
    got = model.get_weights()
    worst = 0.0
    for name in params:
        delta_ref = np.asarray(params[name], dtype=np.float64) - np.asarray(weights[name], dtype=np.float64)
        delta_got = np.asarray(got[name], dtype=np.float64).reshape(delta_ref.shape) - np.asarray(weights[name], dtype=np.float64)
        # Adam's first steps move every coordinate by about lr regardless of gradient scale, so compare the
        # updates in units of lr and allow sign flips only where the gradient is numerically ~0
        err = np.abs(delta_got - delta_ref).mean() / 1e-3
        worst = max(worst, float(err))
    results.append({'name': 'adam mean |update error| / lr (worst tensor)', 'rel': worst, 'tol': 0.35, 'nan': worst != worst,
                    'ok': worst <= 0.35})
    return _finish(results)


def check_generate(B=4, prompt_len=5, length=40, impl=0, max_clusters=0, embedding=256, heads=16, cluster_size=0,
                   window=64, sharp=False):
    '''impl 0 = persistent cluster kernel (decode_mega.cu), 1 = per-step kernels replayed as a CUDA graph.
    ``sharp`` scales the query / key projections so that the attention concentrates on a few cached tokens: with
    the initialiser's weights it is nearly uniform and a k|v row read from the wrong place hardly shows.'''
    _lib.call('cb200_set_decode_impl', impl, max_clusters, cluster_size)
    try:
        return _check_generate(B, prompt_len, length, embedding, heads, window, sharp)
    finally:
        _lib.call('cb200_set_decode_impl', 0, 0, 0)


def check_generate_impls_agree(B=19, prompt_len=3, length=50, embedding=256, heads=16, window=64):
    '''The two decode implementations share every rounding point: same tokens (greedy and sampled) and the
    same final logits up to accumulation order.'''
    import numpy as np

    model, cfg, _ = _small_model(3, embedding, heads, window=window)
    rng = np.random.default_rng(21)
    prompt = rng.integers(0, cfg.vocab_size, size=(B, prompt_len))
    outs = {}
    for impl, clusters in ((1, 0), (0, 0), (0, 2)):
        # the capped run also uses the other cluster size
        _lib.call('cb200_set_decode_impl', impl, clusters, (4 if clusters else 8) if embedding == 256 else 0)
        try:
            greedy, _, logits = model.generate(prompt, length, temperature=0.0, return_uniforms=True, return_last_logits=True)
            sampled = model.generate(prompt, length, temperature=1.0, seed=77)
            outs[(impl, clusters)] = (greedy.cpu().numpy(), sampled.cpu().numpy(), logits.float().cpu())
        finally:
            _lib.call('cb200_set_decode_impl', 0, 0, 0)
    ref = outs[(1, 0)]
    results = []
    for key in ((0, 0), (0, 2)):
        got = outs[key]
        # a near-tie may flip one token and change the continuation: compare up to the first difference
        # (a sampled sequence diverges for good after one draw that lands on the other side of a CDF edge, so only
        # its first tokens are compared; the oracle-based check above replays every step and has no such cascade)
        same_g = float((got[0] == ref[0]).mean())
        same_s = float((got[1][:, :8] == ref[1][:, :8]).mean())
        results.append({'name': 'cluster kernel (clusters %d) greedy tokens equal %.3f' % (key[1], same_g), 'rel': 1 - same_g,
                        'tol': 0.02, 'nan': False, 'ok': same_g >= 0.98})    # measured 1.000 in every variant (profiles/parity_margins_r2.txt)
        results.append({'name': 'cluster kernel (clusters %d) first 8 sampled tokens equal %.3f' % (key[1], same_s),
                        'rel': 1 - same_s, 'tol': 0.12, 'nan': False, 'ok': same_s >= 0.88})    # measured 0.905 .. 1.000: a draw at a CDF edge may fall either side
        rows = (got[0] == ref[0]).all(axis=1)
        if rows.any():
            results.append(_stats('final logits vs per-step kernels (%d identical rows)' % int(rows.sum()),
                                  got[2][torch.from_numpy(rows)], ref[2][torch.from_numpy(rows)], 5e-3))
    a, b = outs[(0, 0)], outs[(0, 2)]
    # the two runs differ in cluster size, hence in the K split of the mlp c_proj: a near-tie may flip a token
    same = float(min((a[1][:, :8] == b[1][:, :8]).mean(), (a[0] == b[0]).mean()))
    results.append({'name': 'tokens independent of the cluster layout %.3f' % same, 'rel': 1 - same, 'tol': 0.02,
                    'nan': False, 'ok': same >= 0.98})
    return _finish(results)


def check_generate_gather_modes_agree(B=13, prompt_len=3, length=150, window=192):
    '''The persistent decode kernel completes its all-gathers either on mbarriers (st.async) or at cluster barriers
    (full clusters; CB200_DECODE_ASYNC_GATHER overrides the launcher's choice).  The arithmetic is the same, so the
    tokens, the uniforms and the final logits have to be identical, for both cluster sizes.'''
    import os
    import numpy as np

    model, cfg, _ = _small_model(3, 256, 16, window=window)
    rng = np.random.default_rng(5)
    prompt = rng.integers(0, cfg.vocab_size, size=(B, prompt_len))
    results = []
    saved = os.environ.get('CB200_DECODE_ASYNC_GATHER')
    try:
        for size in (8, 4):
            outs = []
            for mode in ('0', '1'):
                os.environ['CB200_DECODE_ASYNC_GATHER'] = mode
                _lib.call('cb200_set_decode_impl', 0, 0, size)
                ids, uniforms, logits = model.generate(prompt, length, temperature=1.0, seed=11, return_uniforms=True,
                                                       return_last_logits=True)
                outs.append((ids.cpu().numpy(), uniforms.cpu().numpy(), logits.float().cpu().numpy()))
            same = all(np.array_equal(x, y) for x, y in zip(outs[0], outs[1]))
            results.append({'name': 'cluster size %d: ids, uniforms, logits identical in both gather modes' % size,
                            'rel': 0.0 if same else 1.0, 'tol': 0.0, 'nan': False, 'ok': same})
    finally:
        _lib.call('cb200_set_decode_impl', 0, 0, 0)
        if saved is None:
            os.environ.pop('CB200_DECODE_ASYNC_GATHER', None)
        else:
            os.environ['CB200_DECODE_ASYNC_GATHER'] = saved
    return _finish(results)


def _check_generate(B, prompt_len, length, embedding, heads, window=64, sharp=False):
    import numpy as np
    from oracle import transformer_oracle as oracle

    model, cfg, weights = _small_model(2, embedding, heads, window=window)
    if sharp:
        for name in weights:
            if name.endswith('attn/c_attn/weight'):
                weights[name] = weights[name].copy()
                weights[name][:, :2 * embedding] *= 6.0
        model.set_weights(weights)
    rng = np.random.default_rng(9)
    prompt = rng.integers(0, cfg.vocab_size, size=(B, prompt_len))
    results = []
    # greedy: token-identical wherever the oracle's top-2 margin exceeds the tolerance
    out = model.generate(prompt, length, temperature=0.0).cpu().numpy()
    ids = np.array(prompt)
    mismatches = 0
    checked = 0
    params = oracle.to_torch(weights, torch.float64)
    with torch.no_grad():
        for step in range(length):
            logits, _ = oracle.transformer_call(params, ids, cfg)
            last = logits[:, -1].numpy()
            top2 = np.sort(last, axis=-1)[:, -2:]
            margin = top2[:, 1] - top2[:, 0]
            scale = np.abs(last).max(axis=-1)
            choice = last.argmax(axis=-1)
            decisive = margin > 2e-2 * scale
            checked += int(decisive.sum())
            mismatches += int(((out[:, step] != choice) & decisive).sum())
            # continue from the device's tokens so that a non-decisive divergence does not cascade
            ids = np.concatenate([ids, out[:, step:step + 1]], axis=1)
    results.append({'name': 'greedy decode (decisive steps %d)' % checked, 'rel': mismatches, 'tol': 0, 'nan': False,
                    'ok': mismatches == 0 and checked > 0})
    # sampling: same uniforms through the oracle's inverse CDF
    out_s, uniforms, last_logits = model.generate(prompt, length, temperature=1.0, seed=1234, return_uniforms=True,
                                                  return_last_logits=True)
    out_s = out_s.cpu().numpy()
    uniforms = uniforms.cpu().numpy()
    ids = np.array(prompt)
    agree = 0
    total = 0
    near = 0
    with torch.no_grad():
        for step in range(length):
            logits, _ = oracle.transformer_call(params, ids, cfg)
            p = oracle.next_token_distribution(logits[:, -1].numpy(), 1.0)
            cdf = np.cumsum(p, axis=-1)
            u = uniforms[:, prompt_len - 1 + step]
            chosen = np.minimum((cdf <= u[:, None]).sum(axis=-1), cfg.vocab_size - 1)
            total += B
            agree += int((chosen == out_s[:, step]).sum())
            # a draw within tolerance of a CDF edge may legitimately fall either side
            edge = np.abs(cdf - u[:, None]).min(axis=-1)
            near += int(((chosen != out_s[:, step]) & (edge < 5e-3)).sum())
            if step == length - 1:
                # logits of the last step: every cached token of every layer has been read by then (measured 1-2e-3 of
                # the largest logit in every variant, `tools/parity_margins.py`; a k|v row read from the wrong place
                # in one of two 16-token tiles gave 2.5e-2 with the initialiser's near-uniform attention)
                results.append(_stats('last-step logits vs oracle%s' % (' (sharp attention)' if sharp else ''),
                                      last_logits.float().cpu(), torch.from_numpy(logits[:, -1].numpy()).float(), 8e-3))
            ids = np.concatenate([ids, out_s[:, step:step + 1]], axis=1)
    results.append({'name': 'sampled decode agreement %d/%d (+%d at CDF edges)' % (agree, total, near),
                    'rel': total - agree - near, 'tol': 0, 'nan': False, 'ok': agree + near == total})
    results.append({'name': 'uniforms in [0,1)', 'rel': 0.0, 'tol': 0.0, 'nan': False,
                    'ok': bool((uniforms >= 0).all() and (uniforms < 1).all())})
    # determinism and sharding independence: rows 2.. generated alone with the matching base index
    again = model.generate(prompt[2:], length, temperature=1.0, seed=1234, sequence_index_base=2).cpu().numpy()
    results.append({'name': 'sharding-independent sampling', 'rel': float((again != out_s[2:]).mean()), 'tol': 0.0,
                    'nan': False, 'ok': bool((again == out_s[2:]).all())})
    return _finish(results)


GROUPS['engine'] = [check_engine_forward_backward, check_engine_adam_step,
                    lambda: check_engine_forward_backward(B=1, T=256, layers=3),
                    # the head sizes of the scaled configuration (d_h 64) and of an intermediate one (d_h 32)
                    lambda: check_engine_forward_backward(B=2, T=130, layers=2, embedding=1024, heads=16),
                    lambda: check_engine_forward_backward(B=1, T=200, layers=2, embedding=512, heads=16)]
GROUPS['generate'] = [check_generate, lambda: check_generate(impl=1),
                      lambda: check_generate(B=21, prompt_len=2, length=30, max_clusters=1),
                      lambda: check_generate(B=9, prompt_len=3, length=40, cluster_size=4),
                      lambda: check_generate(B=3, prompt_len=1, length=40, embedding=256, heads=8),
                      lambda: check_generate(B=3, prompt_len=2, length=36, embedding=256, heads=4, cluster_size=4),
                      lambda: check_generate(B=5, prompt_len=4, length=24, embedding=512, heads=8),
                      lambda: check_generate(B=3, prompt_len=4, length=24, embedding=512, heads=16, impl=1),
                      # the scaled configuration's width and head size (d_model 1024, d_h 64: per-step kernels)
                      lambda: check_generate(B=4, prompt_len=4, length=32, embedding=1024, heads=16),
                      # contexts beyond one 64-token chunk of the cluster kernel's KV ring (several chunks per pair,
                      # a partial last chunk, ring stages re-used within a layer), 8- and 4-CTA clusters
                      lambda: check_generate(B=3, prompt_len=70, length=90, window=192),
                      lambda: check_generate(B=10, prompt_len=2, length=140, window=192, cluster_size=4),
                      lambda: check_generate_impls_agree(B=6, prompt_len=2, length=250, window=256),
                      check_generate_gather_modes_agree,
                      # sharp attention (see check_generate): d_h 16 with 8- and 4-CTA clusters and split pairs, d_h 32,
                      # the per-step kernels, and d_h 64
                      lambda: check_generate(B=3, prompt_len=40, length=150, window=192, sharp=True),
                      lambda: check_generate(B=40, prompt_len=2, length=100, window=128, cluster_size=4, sharp=True),
                      lambda: check_generate(B=5, prompt_len=30, length=70, embedding=512, heads=16, window=128, sharp=True),
                      lambda: check_generate(B=3, prompt_len=30, length=70, window=128, impl=1, sharp=True),
                      lambda: check_generate(B=3, prompt_len=20, length=40, embedding=1024, heads=16, sharp=True),
                      check_generate_impls_agree,
                      lambda: check_generate_impls_agree(B=33, embedding=512, heads=16)]


# ---------------------------------------------------------------------------
# BASELINE.json configurations at their stated shapes
# ---------------------------------------------------------------------------

def check_config0_forward_loss(B=4, T=1024, layers=8, embedding=256, heads=16):
    '''configs[0] exactly: default_config.yml hyperparameters (8 blocks, d_model 256, 16 heads, window 1024),
    forward + loss on a batch of 4 sequences; logits <= 2e-2 of the largest logit, loss <= 5e-4 relative.'''
    import numpy as np
    from oracle import transformer_oracle as oracle

    model, cfg, weights = _small_model(layers, embedding, heads, window=T)
    rng = np.random.default_rng(101)
    draw = rng.integers(0, cfg.vocab_size, size=(B, T + 1))
    x, y = draw[:, :-1], draw[:, 1:]
    loss_sum, correct, logits = model.forward_loss(x, y, training=False, return_logits=True)
    torch.cuda.synchronize()
    params = oracle.to_torch(weights, torch.float64)
    with torch.no_grad():
        ref_logits, _ = oracle.transformer_call(params, x, cfg)
        ref_loss = float(oracle.sparse_categorical_crossentropy(y, ref_logits))
        ref_acc = float(oracle.batch_accuracy(y, ref_logits))
    got_loss = float(loss_sum) / (B * T)
    results = [_stats('configs[0] logits (L%d E%d H%d T%d B%d)' % (layers, embedding, heads, T, B),
                      logits.double().cpu().reshape(B * T, -1), ref_logits.reshape(B * T, -1), 2e-2),
               {'name': 'configs[0] loss', 'got': got_loss, 'ref': ref_loss, 'rel': abs(got_loss - ref_loss) / abs(ref_loss),
                'tol': 5e-4, 'nan': got_loss != got_loss, 'ok': abs(got_loss - ref_loss) <= 5e-4 * abs(ref_loss)},
               {'name': 'configs[0] accuracy', 'got': float(correct) / (B * T), 'ref': ref_acc,
                'rel': abs(float(correct) / (B * T) - ref_acc), 'tol': 0.01, 'nan': False,
                'ok': abs(float(correct) / (B * T) - ref_acc) <= 0.01}]
    # the model call (Transformer.__call__) returns the same logits
    again, _ = model(x)
    results.append(_stats('configs[0] __call__ logits', again.double().cpu().reshape(B * T, -1),
                          ref_logits.reshape(B * T, -1), 2e-2))
    return _finish(results)


def check_greedy_sweep(layers=12, embedding=1024, heads=16, prompts=16, prompt_len=8, length=256, window=512, seed=33):
    '''configs[4]: the scaled model (12 blocks, d_model 1024, d_h 64): forward logits on a batch, then a greedy
    decode of `prompts` x `length` steps that must be token-identical with the oracle's cached (`past=`) decode at
    every step where the oracle's top-2 logit margin exceeds the tolerance (2e-2 of the largest logit).  The oracle
    is fed the device's tokens, so that a non-decisive divergence does not cascade.'''
    import numpy as np
    from oracle import transformer_oracle as oracle

    model, cfg, weights = _small_model(layers, embedding, heads, window=window, seed=seed)
    rng = np.random.default_rng(seed + 1)
    params = oracle.to_torch(weights, torch.float64)
    results = []
    # forward logits at the scaled width
    draw = rng.integers(0, cfg.vocab_size, size=(2, 129))
    x = draw[:, :-1]
    logits, _ = model(x)
    with torch.no_grad():
        ref_logits, _ = oracle.transformer_call(params, x, cfg)
    results.append(_stats('scaled model logits (L%d E%d)' % (layers, embedding), logits.double().cpu().reshape(256, -1),
                          ref_logits.reshape(256, -1), 2e-2))
    prompt = rng.integers(0, cfg.vocab_size, size=(prompts, prompt_len))
    out = model.generate(prompt, length, temperature=0.0).cpu().numpy()
    checked = mismatches = 0
    worst_margin_of_mismatch = 0.0
    with torch.no_grad():
        past = None
        context = torch.as_tensor(prompt).long()
        for step in range(length):
            step_logits, past = oracle.transformer_call(params, context, cfg, past=past)
            last = step_logits[:, -1].numpy()
            top2 = np.sort(last, axis=-1)[:, -2:]
            margin = top2[:, 1] - top2[:, 0]
            scale = np.abs(last).max(axis=-1)
            decisive = margin > 2e-2 * scale
            wrong = (out[:, step] != last.argmax(axis=-1)) & decisive
            checked += int(decisive.sum())
            mismatches += int(wrong.sum())
            if wrong.any():
                worst_margin_of_mismatch = max(worst_margin_of_mismatch, float((margin / scale)[wrong].max()))
            context = torch.as_tensor(out[:, step:step + 1]).long()
    results.append({'name': 'greedy sweep %d prompts x %d steps (decisive %d of %d)' % (prompts, length, checked,
                                                                                    prompts * length),
                    'rel': mismatches, 'tol': 0, 'nan': False, 'worst_margin_of_mismatch': worst_margin_of_mismatch,
                    'ok': mismatches == 0 and checked >= prompts * length // 4})
    return _finish(results)


GROUPS['configs'] = [
    check_config0_forward_loss,
    # configs[1]'s sequence length: forward, loss and every gradient at T 2048 (2 blocks keep the CPU oracle's
    # attention matrices, [B, H, T, T] per block, within a few GB; fp32 oracle for the same reason)
    lambda: check_engine_forward_backward(B=2, T=2048, layers=2, oracle_dtype=torch.float32),
    # the block stack without LayerNorm (use_layer_normalization: false; only ln_f remains, transformer.py:583-594, 811)
    lambda: check_engine_forward_backward(B=2, T=100, layers=3, layer_norm=False),
    check_greedy_sweep,
    # the default model through the persistent decode kernel: 16 prompts x 256 steps as well
    lambda: check_greedy_sweep(layers=8, embedding=256, heads=16, prompts=16, prompt_len=4, length=256, window=320),
]


def check_model_call_with_past(B=3, T=37, steps=6, layers=3, embedding=256, heads=16, window=64):
    '''
    The model call of the reference (transformer.py:696-833) through its `past` / `presents` interface:
    ``logits, presents = model(x)`` (batched prefill, cb200_prefill), then ``model(next, past=presents)`` one token
    at a time (cb200_decode_step), against ``oracle.transformer_call`` with the same past.
    '''
    import numpy as np
    from oracle import transformer_oracle as oracle

    model, cfg, weights = _small_model(layers, embedding, heads, window=window)
    params = oracle.to_torch(weights, torch.float64)
    rng = np.random.default_rng(31)
    x = rng.integers(0, cfg.vocab_size, size=(B, T))
    results = []
    logits, presents = model(x)
    with torch.no_grad():
        ref_logits, ref_presents = oracle.transformer_call(params, x, cfg)
    results.append(_stats('prefill logits [B, T, V]', logits.double().cpu().reshape(B * T, -1), ref_logits.reshape(B * T, -1), 2e-2))
    results.append({'name': 'presents: %d layers of [2, B, H, T, d_h]' % len(presents), 'rel': 0.0, 'tol': 0.0, 'nan': False,
                    'ok': len(presents) == layers and tuple(presents[0].shape) == (2, B, heads, T, embedding // heads)})
    for layer in (0, layers - 1):
        ref = torch.stack([ref_presents[layer][0], ref_presents[layer][1]]) if isinstance(ref_presents[layer], (tuple, list)) \
            else ref_presents[layer]
        results.append(_stats('presents[%d] vs oracle' % layer, presents[layer].double().cpu().reshape(-1, embedding // heads),
                              ref.reshape(-1, embedding // heads), 1e-2))
    # plain tensors handed back as `past` (copied into a fresh cache) and the in-place path must agree
    as_tuple = tuple(p.clone() for p in presents)
    nxt = rng.integers(0, cfg.vocab_size, size=(B, 1))
    via_tuple, _ = model(nxt, past=as_tuple)
    past_ref = ref_presents
    for step in range(steps):
        step_logits, presents = model(nxt, past=presents)
        with torch.no_grad():
            ref_step, past_ref = oracle.transformer_call(params, nxt, cfg, past=past_ref)
        if step == 0:
            results.append(_stats('step 0: past given as a tuple of tensors == in-place cache', via_tuple.cpu().reshape(B, -1),
                                  step_logits.cpu().reshape(B, -1), 1e-6))
        if step in (0, steps - 1):
            results.append(_stats('step %d logits [B, 1, V] with past of %d tokens' % (step, T + step),
                                  step_logits.double().cpu().reshape(B, -1), ref_step.reshape(B, -1), 2e-2))
        nxt = rng.integers(0, cfg.vocab_size, size=(B, 1))
    results.append({'name': 'presents grew to %d positions' % presents.length, 'rel': 0.0, 'tol': 0.0, 'nan': False,
                    'ok': presents.length == T + steps and tuple(presents[1].shape)[3] == T + steps})
    # inputs longer than one token with a past: only the last one is used (transformer.py:735-737)
    longer = np.concatenate([rng.integers(0, cfg.vocab_size, size=(B, 4)), nxt], axis=1)
    a, presents = model(longer, past=presents)
    with torch.no_grad():
        b, _ = oracle.transformer_call(params, longer, cfg, past=past_ref)
    results.append(_stats('inputs[:, -1:] is what a call with past uses', a.double().cpu().reshape(B, -1), b.reshape(B, -1), 2e-2))
    return _finish(results)


def check_prefill_equals_teacher_forcing(B=5, prompt_len=23, length=12, embedding=256, heads=16, window=64):
    '''The batched prompt prefill of cb200_generate against feeding the prompt token by token (both decode paths).'''
    import numpy as np

    model, cfg, _ = _small_model(3, embedding, heads, window=window)
    rng = np.random.default_rng(41)
    prompt = rng.integers(0, cfg.vocab_size, size=(B, prompt_len))
    results = []
    for impl in (0, 1):
        outs = {}
        for prefill in (1, 0):
            _lib.call('cb200_set_decode_impl', impl, 0, 0)
            _lib.call('cb200_set_decode_prefill', prefill)
            try:
                ids, uniforms, last = model.generate(prompt, length, temperature=1.0, seed=5, return_uniforms=True,
                                                     return_last_logits=True)
                outs[prefill] = (ids.cpu().numpy(), uniforms.cpu().numpy(), last.float().cpu())
            finally:
                _lib.call('cb200_set_decode_impl', 0, 0, 0)
                _lib.call('cb200_set_decode_prefill', 1)
        same = float((outs[1][0] == outs[0][0]).mean())
        results.append({'name': 'impl %d: sampled tokens, prefill vs teacher forcing, equal %.3f' % (impl, same),
                        'rel': 1 - same, 'tol': 0.1, 'nan': False, 'ok': same >= 0.9})
        results.append({'name': 'impl %d: same uniforms at the sampled steps' % impl, 'rel': 0.0, 'tol': 0.0, 'nan': False,
                        'ok': bool(np.array_equal(outs[1][1][:, prompt_len - 1:], outs[0][1][:, prompt_len - 1:]))})
        rows = (outs[1][0] == outs[0][0]).all(axis=1)
        if rows.any():
            results.append(_stats('impl %d: last-step logits (%d identical rows)' % (impl, int(rows.sum())),
                                  outs[1][2][torch.from_numpy(rows)], outs[0][2][torch.from_numpy(rows)], 8e-3))
    return _finish(results)


GROUPS['past'] = [
    check_model_call_with_past,
    lambda: check_model_call_with_past(B=2, T=64, steps=3, layers=2, embedding=512, heads=8, window=128),
    check_prefill_equals_teacher_forcing,
    lambda: check_prefill_equals_teacher_forcing(B=33, prompt_len=2, length=20),
]


def check_engine_against_reference_golden():
    '''
    The CUDA engine against outputs of the REFERENCE's own ``transformer.py`` (tests/golden/model_golden.npz, case
    'engine': produced by the unmodified reference model under tests/golden/tf_shim.py), without the oracle in
    between: ``Transformer.call`` logits, the losses / accuracies its own ``train()`` loop logged over three Adam
    steps, the norms of its first-step gradients and of its three-step weight updates, greedy ``past=`` decoding.
    '''
    import os
    import sys
    import numpy as np
    from oracle import transformer_oracle as oracle      # initialiser recipe of the case's weights only
    golden_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    if golden_dir not in sys.path:
        sys.path.insert(0, golden_dir)
    import model_cases
    from composer_b200.models.transformer import Transformer

    name = 'engine'
    index = list(model_cases.CASES).index(name)
    case = model_cases.CASES[name]
    cfg = model_cases.case_config(oracle, case)
    weights = model_cases.case_weights(oracle, cfg, 10 + index)
    batches = model_cases.case_batches(case, 10 + index)
    data = np.load(os.path.join(golden_dir, 'model_golden.npz'))
    golden = {k[len(name) + 1:]: data[k] for k in data.files if k.startswith(name + '/')}

    model = Transformer(case['vocab_size'], case['embedding_size'], case['window_size'], case['decoder_layers_count'],
                        case['attention_head_count'], False, 0.0, 0.02, 0.0, 0.0, 1e-5, case['scale'],
                        case['use_layer_normalization'])
    model.set_weights(weights)
    results = []
    x0 = batches[0][0]
    B, T = x0.shape
    logits, presents = model(x0)
    results.append(_stats('logits vs the reference model', logits.double().cpu().reshape(B * T, -1),
                          torch.from_numpy(golden['logits']).double().reshape(B * T, -1), 2e-2))

    # greedy decoding through past= : token-identical wherever the reference's top-2 margin exceeds the tolerance
    steps = golden['decode_ids'].shape[1]
    out = model.generate(x0[:, :case['prompt']], steps, temperature=0.0).cpu().numpy()
    ref_logits, ref_ids = golden['decode_logits'], golden['decode_ids']
    checked = mismatches = 0
    for b in range(B):
        for step in range(steps):
            if step > 0 and out[b, step - 1] != ref_ids[b, step - 1]:
                break                                    # contexts differ from here on
            top2 = np.sort(ref_logits[b, step])[-2:]
            if top2[1] - top2[0] > 2e-2 * np.abs(ref_logits[b, step]).max():
                checked += 1
                mismatches += int(out[b, step] != ref_ids[b, step])
    results.append({'name': 'greedy decode vs the reference model (decisive steps %d)' % checked, 'rel': mismatches, 'tol': 0,
                    'nan': False, 'ok': mismatches == 0 and checked >= steps // 4})

    # the reference's own training loop: losses and accuracies of three Adam steps, first-step gradients, updates
    model.compile(1e-3)
    for step, (x, y) in enumerate(batches):
        loss_sum, correct = model.forward_loss(x, y, training=True, step=step)
        loss, accuracy = float(loss_sum) / x.size, float(correct) / x.size
        results.append({'name': 'train step %d loss %.5f (reference %.5f)' % (step, loss, golden['step_loss'][step]),
                        'rel': abs(loss - golden['step_loss'][step]), 'tol': 5e-4 * (step + 1), 'nan': loss != loss,
                        'ok': abs(loss - golden['step_loss'][step]) <= 5e-4 * (step + 1)})   # (bf16 trajectories drift apart)
        results.append({'name': 'train step %d accuracy' % step, 'rel': abs(accuracy - golden['step_accuracy'][step]),
                        'tol': 2.0 / x.size, 'nan': False, 'ok': abs(accuracy - golden['step_accuracy'][step]) <= 2.0 / x.size})
        world = model.backward()
        if step == 0:
            grads = model.get_gradients()
            worst = 0.0
            for variable, value in grads.items():
                if 'grad_norm/' + variable in golden:
                    ref = float(golden['grad_norm/' + variable])
                    worst = max(worst, abs(float(np.linalg.norm(value)) - ref) / max(ref, 1e-30))
                elif 'grad/' + variable in golden:
                    ref = golden['grad/' + variable]
                    worst = max(worst, float(np.linalg.norm(value.reshape(ref.shape) - ref) / max(np.linalg.norm(ref), 1e-30)))
            results.append({'name': 'first-step gradients vs the reference (worst tensor, relative L2 / norm)', 'rel': worst,
                            'tol': 5e-2, 'nan': worst != worst, 'ok': worst <= 5e-2})
        model.apply_gradients(1e-3, world)
    trained = model.get_weights()
    worst = 0.0
    for variable, value in trained.items():
        delta = np.linalg.norm(value.astype(np.float64) - weights[variable].astype(np.float64))
        if 'update_norm/' + variable in golden:
            ref = float(golden['update_norm/' + variable])
            worst = max(worst, abs(delta - ref) / max(ref, 1e-30))
        elif 'trained/' + variable in golden:
            ref = np.linalg.norm(golden['trained/' + variable].astype(np.float64) - weights[variable].astype(np.float64))
            worst = max(worst, abs(delta - ref) / max(ref, 1e-30))
    results.append({'name': 'size of the three-step Adam update per tensor vs the reference (worst)', 'rel': worst, 'tol': 5e-2,
                    'nan': worst != worst, 'ok': worst <= 5e-2})
    return _finish(results)


GROUPS['golden'] = [check_engine_against_reference_golden]
