// Causal multi-head attention, flash style (the T x T score matrix is never
// written to HBM), forward and backward, for head sizes 16 / 32 / 64.
//
// Reference semantics (composer/models/transformer.py:331-371): S = q k^T,
// S *= rsqrt(d_h) when scale_attention (:345-348, before the mask), causal
// mask S*b - 1e4*(1-b) (:351-354; the masked probabilities underflow to exactly
// 0 in fp32, so they are skipped here), softmax (:360), dropout on the
// probabilities (:361), P v (:367); heads split/merged as :373-395, i.e. head h
// owns columns [h*d_h, (h+1)*d_h) of the q / k / v thirds of c_attn's output.
//
// This file holds the host launchers, the helper kernels of the backward pass
// (delta, dq store), the dropout-mask export used by the parity tests, and the
// round-1 forward kernel (warp-level mma.sync, scores in registers), which is
// kept as an A/B reference behind cb200_set_attention_fwd_impl(1).  The kernels
// that run by default are the tcgen05 / TMEM ones: attention_fwd_tc.cu
// (forward) and attention_tc.cu (backward).  See DESIGN.md "Attention".
#include "attention.h"
#include "mma_sync.cuh"

namespace cb200 {

__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// Byte offset of 16-byte chunk `chunk` of row `row` in a [rows][D] bf16 tile;
// chunks are XOR-swizzled so that ldmatrix (8 rows x 16 B) is conflict-free.
template <int D>
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
    constexpr int CH = D / 8;                    // chunks per row
    constexpr int RPL = (CH >= 8) ? 1 : 8 / CH;  // rows per 128 bytes
    const int sw = (row / RPL) % (CH >= 8 ? 8 : CH);
    return static_cast<uint32_t>(row * (D * 2) + ((chunk ^ sw) << 4));
}

constexpr int ATT_THREADS = 128;
constexpr int ATT_BC = 64;   // keys per inner block of the forward kernel

// A thread's share of a ROWS x D tile copy: which 16-byte chunks it moves every
// time the tile is (re)loaded.  Offsets are computed once, outside the loops.
template <int D, int ROWS>
struct TileCopy {
    static constexpr int CH = D / 8;
    static constexpr int PER_THREAD = (ROWS * CH + ATT_THREADS - 1) / ATT_THREADS;
    uint32_t soff[PER_THREAD];
    int row[PER_THREAD];
    int gcol[PER_THREAD];
    __device__ __forceinline__ void init(int tid) {
#pragma unroll
        for (int c = 0; c < PER_THREAD; ++c) {
            const int idx = tid + c * ATT_THREADS;
            const int r = (idx < ROWS * CH) ? idx / CH : -1;
            row[c] = r;
            gcol[c] = (idx % CH) * 8;
            soff[c] = (r >= 0) ? tile_off<D>(r, idx % CH) : 0;
        }
    }
    // g points at (row 0, column 0) of the head's slice; rows >= row_limit are zero filled
    __device__ __forceinline__ void issue(uint32_t smem_base, const __nv_bfloat16* g, int ld, int row0, int row_limit) const {
#pragma unroll
        for (int c = 0; c < PER_THREAD; ++c) {
            if (row[c] >= 0) {
                const bool ok = (row0 + row[c]) < row_limit;
                const __nv_bfloat16* src = g + static_cast<size_t>(ok ? (row0 + row[c]) : 0) * ld + gcol[c];
                cp_async_16(smem_base + soff[c], src, ok);
            }
        }
    }
};

// ---------------------------------------------------------------------------
// Forward.  One CTA = 64*MW query rows of one (batch, head); warp w owns MW
// groups of 16 rows.  K/V blocks of 64 keys are double buffered with cp.async.
// ---------------------------------------------------------------------------
template <int D, int MW, bool DROP>
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse,
                int T, int H, float scale_log2, AttnDropKey drop) {
    constexpr int BR = 64 * MW;
    constexpr int KS = D / 16;     // k-steps of the QK^T product
    constexpr int NT_O = D / 8;    // n-tiles of the output
    __shared__ __align__(128) uint8_t sQ[BR * D * 2];
    __shared__ __align__(128) uint8_t sK[2][ATT_BC * D * 2];
    __shared__ __align__(128) uint8_t sV[2][ATT_BC * D * 2];

    const int E = H * D;
    const int ld = 3 * E;
    const int qb = gridDim.x - 1 - blockIdx.x;   // heaviest (longest) query blocks first
    const int h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform by construction
    const int g = lane >> 2, tig = lane & 3;
    const int q0 = qb * BR;

    const __nv_bfloat16* base = qkv + static_cast<size_t>(b) * T * ld;
    const __nv_bfloat16* gq = base + h * D;
    const __nv_bfloat16* gk = base + E + h * D;
    const __nv_bfloat16* gv = base + 2 * E + h * D;

    TileCopy<D, ATT_BC> kv_copy;
    kv_copy.init(tid);
    {
        TileCopy<D, BR> q_copy;
        q_copy.init(tid);
        q_copy.issue(smem_u32(sQ), gq, ld, q0, T);
    }
    kv_copy.issue(smem_u32(sK[0]), gk, ld, 0, T);
    kv_copy.issue(smem_u32(sV[0]), gv, ld, 0, T);
    cp_async_commit();

    float o[MW][NT_O][4];
    float m_lo[MW], m_hi[MW], l_lo[MW], l_hi[MW];
    uint32_t qf[MW][KS][4];
#pragma unroll
    for (int mt = 0; mt < MW; ++mt) {
        m_lo[mt] = m_hi[mt] = -INFINITY;
        l_lo[mt] = l_hi[mt] = 0.f;
#pragma unroll
        for (int t = 0; t < NT_O; ++t) { o[mt][t][0] = o[mt][t][1] = o[mt][t][2] = o[mt][t][3] = 0.f; }
    }
    const uint32_t drop_base = DROP ? attn_drop_base(drop, b * H + h) : 0u;
    const float drop_thr = __uint_as_float(drop.thr_bits);
    uint32_t lane_jump_lo = 1u;                      // A^tig and A^(32 + tig)
    for (int i = 0; i < (lane & 3); ++i) lane_jump_lo *= ATTN_MCG_A;
    const uint32_t lane_jump_hi = lane_jump_lo * mcg_mul_pow(32);

    // keys 0 .. min(T, q0 + BR) - 1 are visible to this CTA
    const int kv_end = min(T, q0 + BR);
    const int nblocks = (kv_end + ATT_BC - 1) / ATT_BC;
    // per-lane ldmatrix offsets (do not depend on the block)
    const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_chunk = lane >> 4;               // A fragments (Q)
    const int k_row = ((lane >> 4) << 3) + (lane & 7), k_chunk = (lane >> 3) & 1;            // B fragments of K
    const int v_row = (((lane >> 3) & 1) << 3) + (lane & 7), v_chunk = lane >> 4;            // B fragments of V (trans)

    for (int j = 0; j < nblocks; ++j) {
        const int buf = j & 1;
        if (j + 1 < nblocks) {
            kv_copy.issue(smem_u32(sK[buf ^ 1]), gk, ld, (j + 1) * ATT_BC, T);
            kv_copy.issue(smem_u32(sV[buf ^ 1]), gv, ld, (j + 1) * ATT_BC, T);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (j == 0) {
#pragma unroll
            for (int mt = 0; mt < MW; ++mt)
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
                    ldmatrix_x4(qf[mt][ks], smem_u32(sQ) + tile_off<D>((warp * MW + mt) * 16 + a_row, ks * 2 + a_chunk));
        }
        const uint32_t kbase = smem_u32(sK[buf]), vbase = smem_u32(sV[buf]);
        const int key0 = j * ATT_BC;

#pragma unroll
        for (int mt = 0; mt < MW; ++mt) {
            const int row0 = q0 + (warp * MW + mt) * 16;       // first query row of this 16-row group
            if (key0 <= row0 + 15) {                            // otherwise the whole block is masked for the group
                // ---- S = Q K^T -------------------------------------------------
                float s[8][4];
#pragma unroll
                for (int t = 0; t < 8; ++t) { s[t][0] = s[t][1] = s[t][2] = s[t][3] = 0.f; }
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
                    for (int tp = 0; tp < 4; ++tp) {   // pairs of key n-tiles
                        uint32_t kf[4];
                        ldmatrix_x4(kf, kbase + tile_off<D>(tp * 16 + k_row, ks * 2 + k_chunk));
                        mma_bf16_16816(s[2 * tp], qf[mt][ks], kf[0], kf[1]);
                        mma_bf16_16816(s[2 * tp + 1], qf[mt][ks], kf[2], kf[3]);
                    }
                }
                // ---- causal mask (only blocks that straddle the diagonal) ----------
                if (key0 + ATT_BC - 1 > row0) {
                    const int r_lo = row0 + g, r_hi = r_lo + 8;
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int col = key0 + t * 8 + 2 * tig + e;
                            if (col > r_lo) s[t][e] = -INFINITY;
                            if (col > r_hi) s[t][2 + e] = -INFINITY;
                        }
                    }
                }
                // ---- online softmax ----------------------------------------------
                float bm_lo = m_lo[mt], bm_hi = m_hi[mt];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    bm_lo = fmax3(bm_lo, s[t][0], s[t][1]);
                    bm_hi = fmax3(bm_hi, s[t][2], s[t][3]);
                }
                bm_lo = fmaxf(bm_lo, __shfl_xor_sync(0xffffffffu, bm_lo, 1));
                bm_lo = fmaxf(bm_lo, __shfl_xor_sync(0xffffffffu, bm_lo, 2));
                bm_hi = fmaxf(bm_hi, __shfl_xor_sync(0xffffffffu, bm_hi, 1));
                bm_hi = fmaxf(bm_hi, __shfl_xor_sync(0xffffffffu, bm_hi, 2));
                const float corr_lo = fast_exp2((m_lo[mt] - bm_lo) * scale_log2);
                const float corr_hi = fast_exp2((m_hi[mt] - bm_hi) * scale_log2);
                m_lo[mt] = bm_lo; m_hi[mt] = bm_hi;
                const float ms_lo = bm_lo * scale_log2, ms_hi = bm_hi * scale_log2;
                float ps_lo = 0.f, ps_hi = 0.f;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    s[t][0] = fast_exp2(fmaf(s[t][0], scale_log2, -ms_lo));
                    s[t][1] = fast_exp2(fmaf(s[t][1], scale_log2, -ms_lo));
                    s[t][2] = fast_exp2(fmaf(s[t][2], scale_log2, -ms_hi));
                    s[t][3] = fast_exp2(fmaf(s[t][3], scale_log2, -ms_hi));
                    ps_lo += s[t][0] + s[t][1];
                    ps_hi += s[t][2] + s[t][3];
                }
                l_lo[mt] = l_lo[mt] * corr_lo + ps_lo;
                l_hi[mt] = l_hi[mt] * corr_hi + ps_hi;
#pragma unroll
                for (int t = 0; t < NT_O; ++t) {
                    o[mt][t][0] *= corr_lo; o[mt][t][1] *= corr_lo; o[mt][t][2] *= corr_hi; o[mt][t][3] *= corr_hi;
                }
                // ---- dropout on the probabilities (the row sums above stay undropped) ----
                // The mask is defined per (row, 128-key block) stream (common.cuh); this lane owns pairs
                // 32*(j & 1) + 4 t + tig of rows g and g + 8, i.e. it walks each row's stream in steps of A^4.
                if (DROP) {
                    const uint32_t jb = static_cast<uint32_t>(key0) >> 7;
                    const uint32_t jump = (key0 & 64) ? lane_jump_hi : lane_jump_lo;
                    uint32_t x_lo = attn_row_seed(drop_base, static_cast<uint32_t>(row0 + g), jb) * jump;
                    uint32_t x_hi = attn_row_seed(drop_base, static_cast<uint32_t>(row0 + g + 8), jb) * jump;
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        float m0, m1;
                        uint32_t y = x_lo;
                        attn_drop_pair(y, drop_thr, m0, m1);
                        s[t][0] *= m0; s[t][1] *= m1;
                        y = x_hi;
                        attn_drop_pair(y, drop_thr, m0, m1);
                        s[t][2] *= m0; s[t][3] *= m1;
                        x_lo *= mcg_mul_pow(4);
                        x_hi *= mcg_mul_pow(4);
                    }
                }
                // ---- O += P V -----------------------------------------------------
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {   // 16 keys per step
                    uint32_t pa[4];
                    pa[0] = pack_bf16(s[2 * ks][0], s[2 * ks][1]);
                    pa[1] = pack_bf16(s[2 * ks][2], s[2 * ks][3]);
                    pa[2] = pack_bf16(s[2 * ks + 1][0], s[2 * ks + 1][1]);
                    pa[3] = pack_bf16(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
                    for (int np = 0; np < NT_O / 2; ++np) {
                        uint32_t vf[4];
                        ldmatrix_x4_trans(vf, vbase + tile_off<D>(ks * 16 + v_row, np * 2 + v_chunk));
                        mma_bf16_16816(o[mt][2 * np], pa, vf[0], vf[1]);
                        mma_bf16_16816(o[mt][2 * np + 1], pa, vf[2], vf[3]);
                    }
                }
            }
        }
        __syncthreads();   // everyone is done with buf before it is refilled
    }

    // ---- finalize ---------------------------------------------------------
    const float ks_scale = DROP ? drop.keep_scale : 1.0f;
    __nv_bfloat16* ob = out + static_cast<size_t>(b) * T * E + h * D;
    float* lb = lse + (static_cast<size_t>(b) * H + h) * T;
#pragma unroll
    for (int mt = 0; mt < MW; ++mt) {
        float ll = l_lo[mt], lh = l_hi[mt];
        ll += __shfl_xor_sync(0xffffffffu, ll, 1);
        ll += __shfl_xor_sync(0xffffffffu, ll, 2);
        lh += __shfl_xor_sync(0xffffffffu, lh, 1);
        lh += __shfl_xor_sync(0xffffffffu, lh, 2);
        const float inv_lo = ks_scale / ll, inv_hi = ks_scale / lh;
        const int i_lo = q0 + (warp * MW + mt) * 16 + g, i_hi = i_lo + 8;
#pragma unroll
        for (int t = 0; t < NT_O; ++t) {
            const int col = t * 8 + 2 * tig;
            if (i_lo < T) *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(i_lo) * E + col) = pack_bf16(o[mt][t][0] * inv_lo, o[mt][t][1] * inv_lo);
            if (i_hi < T) *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(i_hi) * E + col) = pack_bf16(o[mt][t][2] * inv_hi, o[mt][t][3] * inv_hi);
        }
        if (lse != nullptr && tig == 0) {
            if (i_lo < T) lb[i_lo] = m_lo[mt] * scale_log2 + log2f(ll);
            if (i_hi < T) lb[i_hi] = m_hi[mt] * scale_log2 + log2f(lh);
        }
    }
}

// ---------------------------------------------------------------------------
// Backward, step 1: delta[b, h, t] = sum_d dO[b, t, h*D + d] * O[b, t, h*D + d]
// ---------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ out,
                      float* __restrict__ delta, int rows, int T, int H) {
    constexpr int LPH = D / 8;   // lanes per head
    const int E = H * D;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int b = row / T, t = row % T;
    for (int c8 = lane; c8 < E / 8; c8 += 32) {
        const uint4 ra = *reinterpret_cast<const uint4*>(dout + static_cast<size_t>(row) * E + c8 * 8);
        const uint4 rb = *reinterpret_cast<const uint4*>(out + static_cast<size_t>(row) * E + c8 * 8);
        const uint32_t wa[4] = {ra.x, ra.y, ra.z, ra.w}, wb[4] = {rb.x, rb.y, rb.z, rb.w};
        float acc = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 x = unpack_bf16(wa[e]), y = unpack_bf16(wb[e]);
            acc += x.x * y.x + x.y * y.y;
        }
#pragma unroll
        for (int o = 1; o < LPH; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((lane % LPH) == 0) {
            const int h = (c8 * 8) / D;
            delta[(static_cast<size_t>(b) * H + h) * T + t] = acc;
        }
    }
}

// dq (fp32 accumulation buffer, [rows, E]) -> bf16 into the q third of dqkv ([rows, 3E]); re-zeroes the buffer.
// One item = 8 columns: two 16-byte loads, two 16-byte zero stores, one 16-byte bf16 store.  A thread keeps 4
// items in flight (the kernel is pure latency otherwise: 1.2 TB/s with one item per thread).
__global__ void __launch_bounds__(256)
attn_dq_store_kernel(float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dqkv, size_t rows, int E) {
    constexpr int U = 4;
    const size_t n8 = rows * (E / 8);
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i0 = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i0 < n8; i0 += U * stride) {
        float4 v0[U], v1[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = i0 + u * stride;
            if (i < n8) {
                const float4* src = reinterpret_cast<const float4*>(dq_acc + i * 8);      // [rows, E] is contiguous
                v0[u] = __ldcs(src);
                v1[u] = __ldcs(src + 1);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = i0 + u * stride;
            if (i < n8) {
                float4* src = reinterpret_cast<float4*>(dq_acc + i * 8);
                src[0] = make_float4(0.f, 0.f, 0.f, 0.f);
                src[1] = make_float4(0.f, 0.f, 0.f, 0.f);
                const size_t row = i / (E / 8);
                const int c = static_cast<int>(i % (E / 8)) * 8;
                uint4 o;
                o.x = pack_bf16(v0[u].x, v0[u].y); o.y = pack_bf16(v0[u].z, v0[u].w);
                o.z = pack_bf16(v1[u].x, v1[u].y); o.w = pack_bf16(v1[u].z, v1[u].w);
                *reinterpret_cast<uint4*>(dqkv + row * 3 * E + c) = o;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Host launchers
// ---------------------------------------------------------------------------
static int g_attention_fwd_impl = 0;   // 0: tcgen05, P in TMEM (TS MMA); 1: round-1 mma.sync kernel; 2: tcgen05, P through smem; 3, 4: tile-shape variants of 0; 7: 0 with two threads per score row
void attention_set_fwd_impl(int impl) { g_attention_fwd_impl = impl; }

static const float kLog2e = 1.4426950408889634f;

template <int D, int MW>
static void launch_fwd(bool dropping, const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse, int T, int H, int B,
                       float c, const AttnDropKey& key, cudaStream_t s) {
    dim3 grid((T + 64 * MW - 1) / (64 * MW), H, B);
    if (dropping) attn_fwd_kernel<D, MW, true><<<grid, ATT_THREADS, 0, s>>>(qkv, out, lse, T, H, c, key);
    else          attn_fwd_kernel<D, MW, false><<<grid, ATT_THREADS, 0, s>>>(qkv, out, lse, T, H, c, key);
}

int attention_fwd(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse, int B, int T, int H, int D, float scale,
                  const DropoutParams& drop, uint32_t layer, cudaStream_t s) {
    if (B * T == 0) return 0;
    CB200_REQUIRE(T <= (1 << 17), "sequences above 2^17 tokens are not supported by the attention dropout stream");
    const AttnDropKey key = make_attn_drop_key(drop, layer);
    if (g_attention_fwd_impl != 1)
        return attention_fwd_tc(qkv, out, lse, B, T, H, D, scale, key, g_attention_fwd_impl, s);
    const float c = scale * kLog2e;
    const bool dropping = key.thr_bits != 0;
    switch (D) {
        case 16: launch_fwd<16, 2>(dropping, qkv, out, lse, T, H, B, c, key, s); break;
        case 32: launch_fwd<32, 1>(dropping, qkv, out, lse, T, H, B, c, key, s); break;
        case 64: launch_fwd<64, 1>(dropping, qkv, out, lse, T, H, B, c, key, s); break;
        default: set_error("attention head size %d is not supported (16, 32 or 64)", D); return -1;
    }
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

int attention_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* out, const __nv_bfloat16* dout, const float* lse,
                  float* delta, float* dq_acc, __nv_bfloat16* dqkv, int B, int T, int H, int D, float scale,
                  const DropoutParams& drop, uint32_t layer, cudaStream_t s) {
    if (B * T == 0) return 0;
    const int rows = B * T;
    const int E = H * D;
    const AttnDropKey key = make_attn_drop_key(drop, layer);
    switch (D) {
        case 16: attn_bwd_delta_kernel<16><<<(rows + 7) / 8, 256, 0, s>>>(dout, out, delta, rows, T, H); break;
        case 32: attn_bwd_delta_kernel<32><<<(rows + 7) / 8, 256, 0, s>>>(dout, out, delta, rows, T, H); break;
        case 64: attn_bwd_delta_kernel<64><<<(rows + 7) / 8, 256, 0, s>>>(dout, out, delta, rows, T, H); break;
        default: set_error("attention head size %d is not supported (16, 32 or 64)", D); return -1;
    }
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    int rc = attention_bwd_tc_main(qkv, dout, lse, delta, dq_acc, dqkv, B, T, H, D, scale, key, s);
    if (rc) return rc;
    size_t n8 = static_cast<size_t>(rows) * (E / 8);
    size_t blocks = (n8 + 4 * 256 - 1) / (4 * 256);
    if (blocks > 8192) blocks = 8192;
    attn_dq_store_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(dq_acc, dqkv, rows, E);
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

// Debug/parity helper: materialise the attention-probability keep mask
// ([B, H, T, T] bytes, 1 = kept) that the kernels apply.  One thread per
// (row, 128-key block) replays that stream.
__global__ void attn_mask_export_kernel(uint8_t* __restrict__ mask, int T, int H, AttnDropKey drop) {
    const int bh = blockIdx.y;
    const int nb = (T + 127) / 128;
    const int total = T * nb;
    const uint32_t base = attn_drop_base(drop, bh);
    const float thr = __uint_as_float(drop.thr_bits);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int i = idx / nb, jb = idx % nb;
        uint32_t x = attn_row_seed(base, i, jb);
        for (int p = 0; p < 64; ++p) {
            float m0, m1;
            attn_drop_pair(x, thr, m0, m1);
            const int j = jb * 128 + 2 * p;
            uint8_t* row = mask + (static_cast<size_t>(bh) * T + i) * T;
            if (j < T) row[j] = (drop.thr_bits == 0 || m0 != 0.f) ? 1 : 0;
            if (j + 1 < T) row[j + 1] = (drop.thr_bits == 0 || m1 != 0.f) ? 1 : 0;
        }
    }
}

int attention_mask_export(uint8_t* mask, int B, int T, int H, const DropoutParams& drop, uint32_t layer, cudaStream_t s) {
    dim3 grid(64, B * H);
    attn_mask_export_kernel<<<grid, 256, 0, s>>>(mask, T, H, make_attn_drop_key(drop, layer));
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

}  // namespace cb200
