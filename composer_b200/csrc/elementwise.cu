// HBM-bound kernels of the training path: embedding gather (+positional add,
// dropout) and its scatter-add backward, LayerNorm forward/backward, bias
// gradients (+dropout backward), TF-style Adam on the flat parameter arena,
// and the bf16 shadow refresh.  All are one-pass, 128-bit vectorised, one
// warp per token row where rows are reduced.
#include "elementwise.h"
#include <cstdlib>

namespace cb200 {

// ---------------------------------------------------------------------------
// Embedding: h = dropout(wte[ids] + wpe[pos0 + t])
// Reference: SharedTokenEmbedding.call(mode='embedding') transformer.py:120-138,
// wpe lookup :786, sum :793, dropout :794.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_fwd_kernel(const int32_t* __restrict__ ids, const float* __restrict__ wte, const float* __restrict__ wpe,
                 __nv_bfloat16* __restrict__ out, int rows, int T, int E, int pos0, int vocab, DropoutParams drop) {
    const int warps_per_block = blockDim.x >> 5;
    const int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    int id = ids[row];
    id = min(max(id, 0), vocab - 1);
    const int pos = pos0 + (row % T);
    const float* te = wte + static_cast<size_t>(id) * E;
    const float* pe = wpe + static_cast<size_t>(pos) * E;
    for (int c8 = lane; c8 < E / 8; c8 += 32) {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(te + c8 * 8));
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(te + c8 * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(pe + c8 * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(pe + c8 * 8 + 4));
        float f[8] = {a0.x + b0.x, a0.y + b0.y, a0.z + b0.z, a0.w + b0.w,
                      a1.x + b1.x, a1.y + b1.y, a1.z + b1.z, a1.w + b1.w};
        if (drop.threshold16 != 0) {
            const Philox4 r = drop_bits_rowmajor(drop, SITE_EMBD, 0, row, c8);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = (drop_u16(r, e) < drop.threshold16) ? 0.f : f[e] * drop.keep_scale;
        }
        uint4 o;
        o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]);
        o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
        *reinterpret_cast<uint4*>(out + static_cast<size_t>(row) * E + c8 * 8) = o;
    }
}

int embed_fwd(const int32_t* ids, const float* wte, const float* wpe, __nv_bfloat16* out, int B, int T, int E,
              int pos0, int vocab, const DropoutParams& drop, cudaStream_t s) {
    CB200_REQUIRE(E % 8 == 0, "embedding size must be a multiple of 8");
    const int rows = B * T;
    if (rows == 0) return 0;
    embed_fwd_kernel<<<(rows + 7) / 8, 256, 0, s>>>(ids, wte, wpe, out, rows, T, E, pos0, vocab, drop);
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

// Backward: dWte[ids[b,t]] += g[b,t], dWpe[pos0+t] += sum_b g[b,t], where
// g = dropout_bwd(dh).  One block per position t; a thread owns 4 columns.
__global__ void __launch_bounds__(256)
embed_bwd_kernel(const int32_t* __restrict__ ids, const __nv_bfloat16* __restrict__ dh, float* __restrict__ dwte,
                 float* __restrict__ dwpe, int B, int T, int E, int pos0, int vocab, DropoutParams drop) {
    const int t = blockIdx.x;
    for (int c4 = threadIdx.x; c4 < E / 4; c4 += blockDim.x) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int b = 0; b < B; ++b) {
            const int row = b * T + t;
            const uint2 raw = *reinterpret_cast<const uint2*>(dh + static_cast<size_t>(row) * E + c4 * 4);
            const float2 lo = unpack_bf16(raw.x), hi = unpack_bf16(raw.y);
            float g[4] = {lo.x, lo.y, hi.x, hi.y};
            if (drop.threshold16 != 0) {
                const Philox4 r = drop_bits_rowmajor(drop, SITE_EMBD, 0, row, c4 >> 1);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    g[e] = (drop_u16(r, (c4 & 1) * 4 + e) < drop.threshold16) ? 0.f : g[e] * drop.keep_scale;
            }
            int id = ids[row];
            id = min(max(id, 0), vocab - 1);
            float* dst = dwte + static_cast<size_t>(id) * E + c4 * 4;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(g[0]), "f"(g[1]), "f"(g[2]),
                         "f"(g[3])
                         : "memory");
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[e] += g[e];
        }
        float* dp = dwpe + static_cast<size_t>(pos0 + t) * E + c4 * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dp), "f"(acc[0]), "f"(acc[1]), "f"(acc[2]),
                     "f"(acc[3])
                     : "memory");
    }
}

int embed_bwd(const int32_t* ids, const __nv_bfloat16* dh, float* dwte, float* dwpe, int B, int T, int E, int pos0,
              int vocab, const DropoutParams& drop, cudaStream_t s) {
    if (B * T == 0) return 0;
    const int threads = (E / 4) < 256 ? ((E / 4 + 31) / 32) * 32 : 256;
    embed_bwd_kernel<<<T, threads, 0, s>>>(ids, dh, dwte, dwpe, B, T, E, pos0, vocab, drop);
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

// ---------------------------------------------------------------------------
// LayerNorm forward (Keras LayerNormalization(epsilon), transformer.py:551,
// 563, 694): biased variance over the last axis, eps inside the rsqrt.
// One warp per row; the row stays in registers between the two passes.
// ---------------------------------------------------------------------------
// A warp normalises RPW rows at a time and issues all their loads before the first reduction: with one 512-byte row
// in flight per warp the kernel was latency-bound at ~3.5 TB/s.
template <int VPL, int RPW>  // 8-element vectors per lane: E = 256 * VPL; rows per warp
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, float* __restrict__ stats,
                     int rows, int E, float eps) {
    const int row0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW;
    if (row0 >= rows) return;
    const int lane = threadIdx.x & 31;
    // The arithmetic runs on packed pairs (FADD2 / FMUL2 / FFMA2): at E = 256 the kernel issued 130 instructions per row
    // and warp at 55 % issue utilisation, as much instruction- as memory-bound.
    uint64_t f[RPW][VPL][4];
    float sum[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int row = row0 + r < rows ? row0 + r : rows - 1;      // (a clamped duplicate keeps the loop uniform)
        uint64_t acc = f2_pack(0.f, 0.f);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const uint4 raw = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * E + (v * 32 + lane) * 8);
            const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 p = unpack_bf16(w[e]);
                f[r][v][e] = f2_pack(p.x, p.y);
                acc = f2_add(acc, f[r][v][e]);
            }
        }
        float lo, hi;
        f2_unpack(acc, lo, hi);
        sum[r] = lo + hi;
    }
    uint64_t g[VPL][4], b[VPL][4];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int c = (v * 32 + lane) * 8;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
        g[v][0] = f2_pack(g0.x, g0.y); g[v][1] = f2_pack(g0.z, g0.w); g[v][2] = f2_pack(g1.x, g1.y); g[v][3] = f2_pack(g1.z, g1.w);
        b[v][0] = f2_pack(b0.x, b0.y); b[v][1] = f2_pack(b0.z, b0.w); b[v][2] = f2_pack(b1.x, b1.y); b[v][3] = f2_pack(b1.z, b1.w);
    }
    const float inv_e = 1.0f / E;
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int row = row0 + r;
        const float mean = warp_sum(sum[r]) * inv_e;
        const uint64_t neg_mean = f2_pack(-mean, -mean);
        uint64_t sq2 = f2_pack(0.f, 0.f);
#pragma unroll
        for (int v = 0; v < VPL; ++v)
#pragma unroll
            for (int e = 0; e < 4; ++e) { const uint64_t d = f2_add(f[r][v][e], neg_mean); sq2 = f2_fma(d, d, sq2); }
        float sq_lo, sq_hi;
        f2_unpack(sq2, sq_lo, sq_hi);
        const float rstd = rsqrtf(warp_sum(sq_lo + sq_hi) * inv_e + eps);
        if (row >= rows) continue;
        if (lane == 0 && stats != nullptr) {
            stats[2 * static_cast<size_t>(row)] = mean;
            stats[2 * static_cast<size_t>(row) + 1] = rstd;
        }
        const uint64_t rstd2 = f2_pack(rstd, rstd);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int c = (v * 32 + lane) * 8;
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                // (x - mean) * rstd * gamma + beta, the rounding points of the scalar expression
                float lo, hi;
                f2_unpack(f2_fma(f2_mul(f2_add(f[r][v][e], neg_mean), rstd2), g[v][e], b[v][e]), lo, hi);
                o[e] = pack_bf16(lo, hi);
            }
            *reinterpret_cast<uint4*>(y + static_cast<size_t>(row) * E + c) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

int layernorm_fwd(const __nv_bfloat16* x, const float* gamma, const float* beta, __nv_bfloat16* y, float* stats,
                  int rows, int E, float eps, cudaStream_t s) {
    CB200_REQUIRE(E % 256 == 0 && E <= 2048, "LayerNorm needs the embedding size to be a multiple of 256 (<= 2048), got %d", E);
    if (rows == 0) return 0;
    // rows per warp: 4 at E = 256, 2 at 512, else 1 (the live row values stay within ~64 registers)
    const int rpw = E == 256 ? 4 : E == 512 ? 2 : 1;
    const int grid = (rows + 8 * rpw - 1) / (8 * rpw);
    switch (E / 256) {
        case 1: layernorm_fwd_kernel<1, 4><<<grid, 256, 0, s>>>(x, gamma, beta, y, stats, rows, E, eps); break;
        case 2: layernorm_fwd_kernel<2, 2><<<grid, 256, 0, s>>>(x, gamma, beta, y, stats, rows, E, eps); break;
        case 3: layernorm_fwd_kernel<3, 1><<<grid, 256, 0, s>>>(x, gamma, beta, y, stats, rows, E, eps); break;
        case 4: layernorm_fwd_kernel<4, 1><<<grid, 256, 0, s>>>(x, gamma, beta, y, stats, rows, E, eps); break;
        case 8: layernorm_fwd_kernel<8, 1><<<grid, 256, 0, s>>>(x, gamma, beta, y, stats, rows, E, eps); break;
        default: set_error("unsupported embedding size %d for LayerNorm", E); return -1;
    }
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

// ---------------------------------------------------------------------------
// LayerNorm backward.  dy = dy_a (+ dy_b); with xhat = (x - mean) * rstd:
//   dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
//   dx_out = dx (+ dres) ; dgamma += sum_rows dy * xhat ; dbeta += sum_rows dy
// Persistent blocks: each warp walks rows with a grid stride and keeps its
// dgamma/dbeta partials in registers; one smem reduction + atomics at the end.
// ---------------------------------------------------------------------------
template <int VPL, int RPI>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy_a, const __nv_bfloat16* __restrict__ dy_b,
                     const __nv_bfloat16* __restrict__ x, const float* __restrict__ stats,
                     const float* __restrict__ gamma, const __nv_bfloat16* __restrict__ dres,
                     __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     int rows, int E, LnBwdTail tail) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps_total = gridDim.x * (blockDim.x >> 5);
    // The arithmetic runs on packed pairs (FADD2 / FMUL2 / FFMA2; pair e = columns 2 e, 2 e + 1 of the lane's 8): the
    // scalar version issued 270 instructions per row and warp at 51 % issue utilisation.
    uint64_t gam[VPL][4], pg[VPL][4], pb[VPL][4];
    uint64_t pt[VPL][4];   // column sums of the tail's g = dropout_bwd(dx)
    const uint64_t zero2 = f2_pack(0.f, 0.f);
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int c = (v * 32 + lane) * 8;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
        gam[v][0] = f2_pack(g0.x, g0.y); gam[v][1] = f2_pack(g0.z, g0.w); gam[v][2] = f2_pack(g1.x, g1.y); gam[v][3] = f2_pack(g1.z, g1.w);
#pragma unroll
        for (int e = 0; e < 4; ++e) { pg[v][e] = zero2; pb[v][e] = zero2; pt[v][e] = zero2; }
    }
    const float inv_e = 1.0f / E;
    // RPI rows per warp and iteration, all their loads issued before the first use.  (Measured at E = 256: two rows in
    // flight per warp, 118 registers, 31.7 us against 30.1 us with one: the kernel is not short of bytes in flight;
    // RPI = 1 everywhere.)
    for (int row0 = blockIdx.x * (blockDim.x >> 5) + warp; row0 < rows; row0 += RPI * warps_total) {
        uint4 ra[RPI][VPL], rx[RPI][VPL], rb[RPI][VPL], rr[RPI][VPL];
        float mean[RPI], rstd[RPI];
#pragma unroll
        for (int r = 0; r < RPI; ++r) {
            const int row = row0 + r * warps_total;
            const bool live = row < rows;
            mean[r] = live ? stats[2 * static_cast<size_t>(row)] : 0.f;
            rstd[r] = live ? stats[2 * static_cast<size_t>(row) + 1] : 0.f;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const size_t off = static_cast<size_t>(row) * E + (v * 32 + lane) * 8;
                ra[r][v] = rx[r][v] = rb[r][v] = rr[r][v] = make_uint4(0, 0, 0, 0);
                if (live) {
                    ra[r][v] = *reinterpret_cast<const uint4*>(dy_a + off);
                    rx[r][v] = *reinterpret_cast<const uint4*>(x + off);
                    if (dy_b != nullptr) rb[r][v] = *reinterpret_cast<const uint4*>(dy_b + off);
                    if (RPI > 1 && dres != nullptr) rr[r][v] = *reinterpret_cast<const uint4*>(dres + off);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RPI; ++r) {
            const int row = row0 + r * warps_total;
            if (row >= rows) break;                   // (warp-uniform)
            uint64_t dyv[VPL][4], xh[VPL][4];
            uint64_t s1p = zero2, s2p = zero2;
            const uint64_t neg_mean = f2_pack(-mean[r], -mean[r]), rstd2 = f2_pack(rstd[r], rstd[r]);
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const uint32_t wa[4] = {ra[r][v].x, ra[r][v].y, ra[r][v].z, ra[r][v].w};
                const uint32_t wx[4] = {rx[r][v].x, rx[r][v].y, rx[r][v].z, rx[r][v].w};
                const uint32_t wb[4] = {rb[r][v].x, rb[r][v].y, rb[r][v].z, rb[r][v].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 a = unpack_bf16(wa[e]), b = unpack_bf16(wb[e]), xx = unpack_bf16(wx[e]);
                    dyv[v][e] = f2_add(f2_pack(a.x, a.y), f2_pack(b.x, b.y));
                    xh[v][e] = f2_mul(f2_add(f2_pack(xx.x, xx.y), neg_mean), rstd2);
                    const uint64_t g = f2_mul(dyv[v][e], gam[v][e]);
                    s1p = f2_add(s1p, g);
                    s2p = f2_fma(g, xh[v][e], s2p);
                    pg[v][e] = f2_fma(dyv[v][e], xh[v][e], pg[v][e]);
                    pb[v][e] = f2_add(pb[v][e], dyv[v][e]);
                }
            }
            float s1lo, s1hi, s2lo, s2hi;
            f2_unpack(s1p, s1lo, s1hi);
            f2_unpack(s2p, s2lo, s2hi);
            const float s1 = warp_sum(s1lo + s1hi) * inv_e;
            const float s2 = warp_sum(s2lo + s2hi) * inv_e;
            const uint64_t neg_s1 = f2_pack(-s1, -s1), neg_s2 = f2_pack(-s2, -s2);
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const size_t off = static_cast<size_t>(row) * E + (v * 32 + lane) * 8;
                if (RPI == 1 && dres != nullptr) rr[r][v] = *reinterpret_cast<const uint4*>(dres + off);      // (late: fewer live registers)
                const uint32_t wr[4] = {rr[r][v].x, rr[r][v].y, rr[r][v].z, rr[r][v].w};
                uint32_t wo[4];
                float f[8];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    // rstd * (dy * gamma - s1 - xhat * s2) (+ dres)
                    uint64_t o = f2_mul(rstd2, f2_fma(xh[v][e], neg_s2, f2_fma(dyv[v][e], gam[v][e], neg_s1)));
                    if (dres != nullptr) { const float2 q = unpack_bf16(wr[e]); o = f2_add(o, f2_pack(q.x, q.y)); }
                    float lo, hi;
                    f2_unpack(o, lo, hi);
                    wo[e] = pack_bf16(lo, hi);
                }
                *reinterpret_cast<uint4*>(dx + off) = make_uint4(wo[0], wo[1], wo[2], wo[3]);
                if (tail.dbias != nullptr) {
                    // what bias_grad would do with dx as its input: the bf16-rounded values, the same keep mask
#pragma unroll
                    for (int e = 0; e < 4; ++e) { const float2 q = unpack_bf16(wo[e]); f[2 * e] = q.x; f[2 * e + 1] = q.y; }
                    if (tail.drop.threshold16 != 0) {
                        const Philox4 ph = drop_bits_rowmajor(tail.drop, tail.site, tail.layer, row, v * 32 + lane);
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] = (drop_u16(ph, e) < tail.drop.threshold16) ? 0.f : f[e] * tail.drop.keep_scale;
                        uint4 g;
                        g.x = pack_bf16(f[0], f[1]); g.y = pack_bf16(f[2], f[3]);
                        g.z = pack_bf16(f[4], f[5]); g.w = pack_bf16(f[6], f[7]);
                        *reinterpret_cast<uint4*>(tail.g_out + off) = g;
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) pt[v][e] = f2_add(pt[v][e], f2_pack(f[2 * e], f[2 * e + 1]));
                }
            }
        }
    }
    // cross-warp reduction of the dgamma / dbeta (/ tail bias) partials through shared memory
    extern __shared__ float red[];   // [3][E]
    for (int i = threadIdx.x; i < 3 * E; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = (v * 32 + lane) * 8 + 2 * e;
            float lo, hi;
            f2_unpack(pg[v][e], lo, hi);
            atomicAdd(&red[c], lo); atomicAdd(&red[c + 1], hi);
            f2_unpack(pb[v][e], lo, hi);
            atomicAdd(&red[E + c], lo); atomicAdd(&red[E + c + 1], hi);
            if (tail.dbias != nullptr) {
                f2_unpack(pt[v][e], lo, hi);
                atomicAdd(&red[2 * E + c], lo); atomicAdd(&red[2 * E + c + 1], hi);
            }
        }
    __syncthreads();
    for (int i = threadIdx.x; i < E; i += blockDim.x) {
        atomicAdd(&dgamma[i], red[i]);
        atomicAdd(&dbeta[i], red[E + i]);
        if (tail.dbias != nullptr) atomicAdd(&tail.dbias[i], red[2 * E + i]);
    }
}

// The same for wide rows (E = 256 WPR): WPR warps share a row, each owning 256 columns, so that the per-column
// partial sums stay at 8 per lane and kind (24 registers instead of 96 at E = 1024, where the one-warp-per-row kernel
// ran at 1 block per SM and 1.8 TB/s).  The warps of a row exchange their two partial row sums through shared memory
// behind a named barrier; the exchange slots alternate between rows, so one barrier per row is enough.
template <int WPR>
__global__ void __launch_bounds__(256)
layernorm_bwd_split_kernel(const __nv_bfloat16* __restrict__ dy_a, const __nv_bfloat16* __restrict__ dy_b,
                           const __nv_bfloat16* __restrict__ x, const float* __restrict__ stats,
                           const float* __restrict__ gamma, const __nv_bfloat16* __restrict__ dres,
                           __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                           int rows, int E, LnBwdTail tail) {
    constexpr int GROUPS = 8 / WPR;
    __shared__ float part[GROUPS][2][WPR][2];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int group = warp / WPR, w = warp % WPR;
    const int col = (w * 32 + lane) * 8;              // this lane's 8 columns
    float gam[8], pg[8], pb[8], pt[8];
    {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
        gam[0] = g0.x; gam[1] = g0.y; gam[2] = g0.z; gam[3] = g0.w; gam[4] = g1.x; gam[5] = g1.y; gam[6] = g1.z; gam[7] = g1.w;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) { pg[e] = 0.f; pb[e] = 0.f; pt[e] = 0.f; }
    int buf = 0;
    const float inv_e = 1.0f / E;
    for (int row = blockIdx.x * GROUPS + group; row < rows; row += gridDim.x * GROUPS) {
        const float mean = stats[2 * static_cast<size_t>(row)], rstd = stats[2 * static_cast<size_t>(row) + 1];
        const size_t off = static_cast<size_t>(row) * E + col;
        const uint4 ra = *reinterpret_cast<const uint4*>(dy_a + off);
        const uint4 rx = *reinterpret_cast<const uint4*>(x + off);
        uint4 rb = make_uint4(0, 0, 0, 0);
        if (dy_b != nullptr) rb = *reinterpret_cast<const uint4*>(dy_b + off);
        uint4 rr = make_uint4(0, 0, 0, 0);
        if (dres != nullptr) rr = *reinterpret_cast<const uint4*>(dres + off);
        const uint32_t wa[4] = {ra.x, ra.y, ra.z, ra.w}, wx[4] = {rx.x, rx.y, rx.z, rx.w}, wb[4] = {rb.x, rb.y, rb.z, rb.w};
        float dyv[8], xh[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 a = unpack_bf16(wa[e]), b = unpack_bf16(wb[e]), xx = unpack_bf16(wx[e]);
            dyv[2 * e] = a.x + b.x; dyv[2 * e + 1] = a.y + b.y;
            xh[2 * e] = (xx.x - mean) * rstd; xh[2 * e + 1] = (xx.y - mean) * rstd;
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float g = dyv[e] * gam[e];
            s1 += g; s2 += g * xh[e];
            pg[e] += dyv[e] * xh[e];
            pb[e] += dyv[e];
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0) { part[group][buf][w][0] = s1; part[group][buf][w][1] = s2; }
        asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(32 * WPR) : "memory");
        s1 = 0.f; s2 = 0.f;
#pragma unroll
        for (int k = 0; k < WPR; ++k) { s1 += part[group][buf][k][0]; s2 += part[group][buf][k][1]; }
        buf ^= 1;
        s1 *= inv_e; s2 *= inv_e;
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = rstd * (dyv[e] * gam[e] - s1 - xh[e] * s2);
        if (dres != nullptr) {
            const uint32_t wr[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) { const float2 r = unpack_bf16(wr[e]); o[2 * e] += r.x; o[2 * e + 1] += r.y; }
        }
        uint4 out;
        out.x = pack_bf16(o[0], o[1]); out.y = pack_bf16(o[2], o[3]);
        out.z = pack_bf16(o[4], o[5]); out.w = pack_bf16(o[6], o[7]);
        *reinterpret_cast<uint4*>(dx + off) = out;
        if (tail.dbias != nullptr) {
            const uint32_t wo[4] = {out.x, out.y, out.z, out.w};
            float f[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) { const float2 p = unpack_bf16(wo[e]); f[2 * e] = p.x; f[2 * e + 1] = p.y; }
            if (tail.drop.threshold16 != 0) {
                const Philox4 r = drop_bits_rowmajor(tail.drop, tail.site, tail.layer, row, w * 32 + lane);
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = (drop_u16(r, e) < tail.drop.threshold16) ? 0.f : f[e] * tail.drop.keep_scale;
                uint4 g;
                g.x = pack_bf16(f[0], f[1]); g.y = pack_bf16(f[2], f[3]);
                g.z = pack_bf16(f[4], f[5]); g.w = pack_bf16(f[6], f[7]);
                *reinterpret_cast<uint4*>(tail.g_out + off) = g;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) pt[e] += f[e];
        }
    }
    extern __shared__ float red[];   // [3][E]
    for (int i = threadIdx.x; i < 3 * E; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        atomicAdd(&red[col + e], pg[e]);
        atomicAdd(&red[E + col + e], pb[e]);
        if (tail.dbias != nullptr) atomicAdd(&red[2 * E + col + e], pt[e]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < E; i += blockDim.x) {
        atomicAdd(&dgamma[i], red[i]);
        atomicAdd(&dbeta[i], red[E + i]);
        if (tail.dbias != nullptr) atomicAdd(&tail.dbias[i], red[2 * E + i]);
    }
}

int layernorm_bwd(const __nv_bfloat16* dy_a, const __nv_bfloat16* dy_b, const __nv_bfloat16* x, const float* stats,
                  const float* gamma, const __nv_bfloat16* dres, __nv_bfloat16* dx, float* dgamma, float* dbeta,
                  int rows, int E, cudaStream_t s) {
    LnBwdTail none{};
    return layernorm_bwd_tail(dy_a, dy_b, x, stats, gamma, dres, dx, dgamma, dbeta, rows, E, none, s);
}

int layernorm_bwd_tail(const __nv_bfloat16* dy_a, const __nv_bfloat16* dy_b, const __nv_bfloat16* x, const float* stats,
                       const float* gamma, const __nv_bfloat16* dres, __nv_bfloat16* dx, float* dgamma, float* dbeta,
                       int rows, int E, const LnBwdTail& tail, cudaStream_t s) {
    CB200_REQUIRE(tail.dbias == nullptr || tail.drop.threshold16 == 0 || tail.g_out != nullptr,
                  "the fused dropout-backward tail needs an output buffer");
    CB200_REQUIRE(E % 256 == 0 && E <= 1024, "LayerNorm backward needs E %% 256 == 0 and E <= 1024, got %d", E);
    if (rows == 0) return 0;
    int grid = (rows + 7) / 8;
    // persistent blocks, as many as are resident: 3 per SM at E = 256 (80 registers; 26.0 -> 24.4 us against 2), else 2
    int ln_per_sm = E == 256 ? 3 : 2;
    if (const char* env = getenv("CB200_LN_BWD_BLOCKS_PER_SM")) ln_per_sm = atoi(env) > 0 ? atoi(env) : ln_per_sm;      // tuning knob
    const int cap = ln_per_sm * device_sm_count_ew();
    if (grid > cap) grid = cap;
    const size_t smem = 3 * E * sizeof(float);
    if (E == 1024) {
        int wide = (rows + 1) / 2;                  // 2 rows per block pass
        const int wide_cap = 4 * device_sm_count_ew();
        if (wide > wide_cap) wide = wide_cap;
        layernorm_bwd_split_kernel<4><<<wide, 256, smem, s>>>(dy_a, dy_b, x, stats, gamma, dres, dx, dgamma, dbeta, rows, E, tail);
        CB200_CUDA_OK(cudaGetLastError());
        note_launch(1);
        return 0;
    }
    switch (E / 256) {
        case 1: layernorm_bwd_kernel<1, 1><<<grid, 256, smem, s>>>(dy_a, dy_b, x, stats, gamma, dres, dx, dgamma, dbeta, rows, E, tail); break;
        case 2: layernorm_bwd_kernel<2, 1><<<grid, 256, smem, s>>>(dy_a, dy_b, x, stats, gamma, dres, dx, dgamma, dbeta, rows, E, tail); break;
        case 3: layernorm_bwd_kernel<3, 1><<<grid, 256, smem, s>>>(dy_a, dy_b, x, stats, gamma, dres, dx, dgamma, dbeta, rows, E, tail); break;
        case 4: layernorm_bwd_kernel<4, 1><<<grid, 256, smem, s>>>(dy_a, dy_b, x, stats, gamma, dres, dx, dgamma, dbeta, rows, E, tail); break;
        default: set_error("unsupported embedding size %d for LayerNorm backward", E); return -1;
    }
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

// ---------------------------------------------------------------------------
// Bias gradient (+ dropout backward):  g = dropout_bwd(dy);  dbias += colsum(g);
// optionally writes g (the dgrad/wgrad operand) when dropout is active.
// ---------------------------------------------------------------------------
// A block owns `rows_per_block` rows of one slice of at most 256 column groups (8 columns each);
// blockDim is a multiple of the slice's group count, so a thread keeps one column group for all its rows
// and accumulates in registers; one shared-memory reduction and N atomics per block at the end.
__global__ void __launch_bounds__(256)
bias_grad_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ g_out, float* __restrict__ dbias,
                 int rows, int N, int rows_per_block, int groups_per_slice, DropoutParams drop, uint32_t site,
                 uint32_t layer) {
    extern __shared__ float red[];   // [8 * groups in this slice]
    const int group0 = blockIdx.y * groups_per_slice;
    const int groups = min(groups_per_slice, N / 8 - group0);
    for (int i = threadIdx.x; i < groups * 8; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    const int lanes = blockDim.x / groups;          // rows handled per pass
    const int my_group = threadIdx.x % groups;
    const int my_lane = threadIdx.x / groups;
    const int r0 = blockIdx.x * rows_per_block;
    const int r1 = min(rows, r0 + rows_per_block);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (my_lane < lanes) {
        const int grp = group0 + my_group;
        // 4 rows per pass: the loads of a pass are issued together (one load in flight per thread left the kernel
        // latency-bound at ~3.5 TB/s)
        constexpr int U = 4;
        for (int row0 = r0 + my_lane; row0 < r1; row0 += U * lanes) {
            uint4 raw[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int row = row0 + u * lanes;
                raw[u] = make_uint4(0, 0, 0, 0);
                if (row < r1) raw[u] = *reinterpret_cast<const uint4*>(dy + static_cast<size_t>(row) * N + grp * 8);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int row = row0 + u * lanes;
                if (row >= r1) break;
                const size_t off = static_cast<size_t>(row) * N + grp * 8;
                const uint32_t w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
                float f[8];
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float2 p = unpack_bf16(w[e]); f[2 * e] = p.x; f[2 * e + 1] = p.y; }
                if (drop.threshold16 != 0) {
                    const Philox4 r = drop_bits_rowmajor(drop, site, layer, row, grp);
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = (drop_u16(r, e) < drop.threshold16) ? 0.f : f[e] * drop.keep_scale;
                    if (g_out != nullptr) {
                        uint4 o;
                        o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]);
                        o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
                        *reinterpret_cast<uint4*>(g_out + off) = o;
                    }
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] += f[e];
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(&red[my_group * 8 + e], acc[e]);
    }
    __syncthreads();
    if (dbias != nullptr)
        for (int i = threadIdx.x; i < groups * 8; i += blockDim.x) atomicAdd(&dbias[group0 * 8 + i], red[i]);
}

int bias_grad(const __nv_bfloat16* dy, __nv_bfloat16* g_out, float* dbias, int rows, int N, const DropoutParams& drop,
              uint32_t site, uint32_t layer, cudaStream_t s) {
    CB200_REQUIRE(N % 8 == 0, "bias_grad needs N %% 8 == 0");
    if (rows == 0) return 0;
    const int groups_total = N / 8;
    // slices of at most 256 groups whose size divides evenly: 256, or all groups when there are fewer
    const int groups_per_slice = groups_total <= 256 ? groups_total : 256;
    const int slices = (groups_total + groups_per_slice - 1) / groups_per_slice;
    // the last slice may be smaller; the block size must be a multiple of every slice's group count
    const int last = groups_total - (slices - 1) * groups_per_slice;
    int threads = (256 / groups_per_slice) * groups_per_slice;
    if (slices > 1 && threads % last != 0) threads = 256;     // 256 groups per full slice: any divisor of 256 works
    CB200_REQUIRE(threads % last == 0 || slices == 1, "bias_grad: unsupported width %d", N);
    // one wave of resident blocks (3 per SM at 76 registers): every block ends in N atomics on the same N addresses,
    // and with 8 blocks per SM (2.6 waves) that tail cost a fifth of the time (34.6 -> 27.9 us at 65,536 x 1024)
    int per_sm = 3;
    if (const char* env = getenv("CB200_BIAS_GRAD_BLOCKS_PER_SM")) per_sm = atoi(env) > 0 ? atoi(env) : 3;      // tuning knob
    const int target_blocks = per_sm * device_sm_count_ew() / slices;
    int rows_per_block = (rows + target_blocks - 1) / target_blocks;
    if (rows_per_block < 8) rows_per_block = 8;
    dim3 grid((rows + rows_per_block - 1) / rows_per_block, slices);
    bias_grad_kernel<<<grid, threads, groups_per_slice * 8 * sizeof(float), s>>>(dy, g_out, dbias, rows, N, rows_per_block,
                                                                                  groups_per_slice, drop, site, layer);
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

// ---------------------------------------------------------------------------
// Adam with TF-2 Keras semantics (transformer.py:887, 921):
//   m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2 ; p <- p - lr_t * m / (sqrt(v) + eps)
// lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t) is computed on the host in double.
// Also refreshes the bf16 shadow (same flat layout) used by the GEMMs.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            __nv_bfloat16* __restrict__ shadow, size_t n4, float lr_t, float b1, float b2, float eps,
            float grad_scale) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        float* pa = reinterpret_cast<float*>(&pp);
        const float* ga = reinterpret_cast<const float*>(&gg);
        float* ma = reinterpret_cast<float*>(&mm);
        float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float gr = ga[e] * grad_scale;
            ma[e] = b1 * ma[e] + (1.0f - b1) * gr;
            va[e] = b2 * va[e] + (1.0f - b2) * gr * gr;
            pa[e] = pa[e] - lr_t * ma[e] / (sqrtf(va[e]) + eps);
        }
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
        if (shadow != nullptr) {
            uint2 o;
            o.x = pack_bf16(pa[0], pa[1]); o.y = pack_bf16(pa[2], pa[3]);
            reinterpret_cast<uint2*>(shadow)[i] = o;
        }
    }
}

int adam_step(float* p, const float* g, float* m, float* v, __nv_bfloat16* shadow, size_t n, float lr_t, float b1,
              float b2, float eps, float grad_scale, cudaStream_t s) {
    CB200_REQUIRE(n % 4 == 0, "parameter arena length must be a multiple of 4");
    if (n == 0) return 0;
    const size_t n4 = n / 4;
    size_t blocks = (n4 + 255) / 256;
    const size_t cap = 8 * static_cast<size_t>(device_sm_count_ew());
    if (blocks > cap) blocks = cap;
    adam_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(p, g, m, v, shadow, n4, lr_t, b1, b2, eps, grad_scale);
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

// fp32 -> bf16 cast of a flat array (initial shadow fill after loading parameters).
__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
        dst[i] = __float2bfloat16_rn(src[i]);
}

int cast_bf16(const float* src, __nv_bfloat16* dst, size_t n, cudaStream_t s) {
    if (n == 0) return 0;
    size_t blocks = (n + 255) / 256;
    if (blocks > 2048) blocks = 2048;
    cast_bf16_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(src, dst, n);
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

// Batched transpose-cast: for each job, dst[c, r] (ld = dst_ld) = bf16(src[r, c]).
// Produces the [out, in] ("K-major B operand") shadows of the [in, out] weights.
__global__ void __launch_bounds__(256)
transpose_cast_kernel(const float* __restrict__ params, __nv_bfloat16* __restrict__ dst_base,
                      const TransposeJob* __restrict__ jobs) {
    // 64 x 64 tiles: 256-byte row segments in, 128-byte row segments out (16-byte stores of 8 bf16 each); the 32 x 32
    // version wrote 64 bytes per warp instruction and ran at 2 TB/s on the 152 M parameters of the scaled config.
    __shared__ float tile[64][65];
    const TransposeJob job = jobs[blockIdx.z];
    const int tiles_c = (job.cols + 63) / 64;
    const int tiles_r = (job.rows + 63) / 64;
    const bool vec_in = (job.cols % 4 == 0) && (job.src_offset % 4 == 0);
    const bool vec_out = (job.dst_ld % 8 == 0) && (job.dst_offset % 8 == 0);
    for (int tile_id = blockIdx.x; tile_id < tiles_c * tiles_r; tile_id += gridDim.x) {
        const int tr = tile_id / tiles_c, tc = tile_id % tiles_c;
        __syncthreads();
        {   // load: 16 threads x float4 per row, 16 rows per pass
            const int lc = (threadIdx.x & 15) * 4, lr = threadIdx.x >> 4;
            for (int i = lr; i < 64; i += 16) {
                const int r = tr * 64 + i, c = tc * 64 + lc;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < job.rows) {
                    const float* src = params + job.src_offset + static_cast<size_t>(r) * job.cols + c;
                    if (vec_in && c + 3 < job.cols) v = *reinterpret_cast<const float4*>(src);
                    else {
                        if (c < job.cols) v.x = src[0];
                        if (c + 1 < job.cols) v.y = src[1];
                        if (c + 2 < job.cols) v.z = src[2];
                        if (c + 3 < job.cols) v.w = src[3];
                    }
                }
                tile[i][lc] = v.x; tile[i][lc + 1] = v.y; tile[i][lc + 2] = v.z; tile[i][lc + 3] = v.w;
            }
        }
        __syncthreads();
        {   // store: 8 threads x 8 bf16 per output row (= source column), 32 output rows per pass
            const int sr = (threadIdx.x & 7) * 8, sc = threadIdx.x >> 3;
            for (int i = sc; i < 64; i += 32) {
                const int c = tc * 64 + i, r = tr * 64 + sr;
                if (c >= job.cols || r >= job.rows) continue;
                __nv_bfloat16* dst = dst_base + job.dst_offset + static_cast<size_t>(c) * job.dst_ld + r;
                if (vec_out && r + 7 < job.rows) {
                    uint4 o;
                    o.x = pack_bf16(tile[sr][i], tile[sr + 1][i]); o.y = pack_bf16(tile[sr + 2][i], tile[sr + 3][i]);
                    o.z = pack_bf16(tile[sr + 4][i], tile[sr + 5][i]); o.w = pack_bf16(tile[sr + 6][i], tile[sr + 7][i]);
                    *reinterpret_cast<uint4*>(dst) = o;
                } else {
                    for (int k = 0; k < 8 && r + k < job.rows; ++k) dst[k] = __float2bfloat16_rn(tile[sr + k][i]);
                }
            }
        }
    }
}

int transpose_cast(const float* params, __nv_bfloat16* dst_base, const TransposeJob* jobs_dev, int num_jobs,
                   cudaStream_t s) {
    if (num_jobs == 0) return 0;
    dim3 grid(64, 1, num_jobs);
    transpose_cast_kernel<<<grid, 256, 0, s>>>(params, dst_base, jobs_dev);
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

int device_sm_count_ew() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

}  // namespace cb200
