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
// At the default d_h = 16 one 64x64 score block costs 16 m16n8k16 MMAs but
// 4096 exponentials, so the kernel is bound by MUFU.EX2 and issue slots, not
// by the tensor pipe: scores are produced with warp-level mma.sync into
// registers (the layout the softmax needs) instead of a TMEM round trip.
// See DESIGN.md "Attention" for the arithmetic.
#include "attention.h"

namespace cb200 {

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
    uint32_t d;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
    return d;
}

__device__ __forceinline__ void cp_async_16(uint32_t smem_addr, const void* gmem, bool valid) {
    const int src_bytes = valid ? 16 : 0;   // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Byte offset of 16-byte chunk `chunk` of row `row` in a [rows][D] bf16 tile;
// chunks are XOR-swizzled so that ldmatrix (8 rows x 16 B) is conflict-free.
template <int D>
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
    constexpr int CH = D / 8;                    // chunks per row
    constexpr int RPL = (CH >= 8) ? 1 : 8 / CH;  // rows per 128 bytes
    const int sw = (row / RPL) % (CH >= 8 ? 8 : CH);
    return static_cast<uint32_t>(row * (D * 2) + ((chunk ^ sw) << 4));
}

// Cooperative async load of `rows` x D bf16 (global row stride ld elements) into a swizzled tile.
template <int D, int THREADS>
__device__ __forceinline__ void load_tile_async(uint32_t smem_base, const __nv_bfloat16* g, int ld, int row0,
                                                int rows, int row_limit, int tid) {
    constexpr int CH = D / 8;
    for (int idx = tid; idx < rows * CH; idx += THREADS) {
        const int r = idx / CH, c = idx % CH;
        const bool ok = (row0 + r) < row_limit;
        const __nv_bfloat16* src = g + static_cast<size_t>(ok ? (row0 + r) : 0) * ld + c * 8;
        cp_async_16(smem_base + tile_off<D>(r, c), src, ok);
    }
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

constexpr int ATT_BR = 64;   // query rows per CTA (4 warps x 16)
constexpr int ATT_BC = 64;   // keys per inner block
constexpr int ATT_THREADS = 128;

// ---------------------------------------------------------------------------
// Forward
// ---------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse,
                int T, int H, float scale_log2, AttnDropKey drop) {
    constexpr int KS = D / 16;     // k-steps of the QK^T product
    constexpr int NT_O = D / 8;    // n-tiles of the output
    __shared__ __align__(128) uint8_t sQ[ATT_BR * D * 2];
    __shared__ __align__(128) uint8_t sK[2][ATT_BC * D * 2];
    __shared__ __align__(128) uint8_t sV[2][ATT_BC * D * 2];

    const int E = H * D;
    const int ld = 3 * E;
    const int qb = gridDim.x - 1 - blockIdx.x;   // heaviest (longest) query blocks first
    const int h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tig = lane & 3;
    const int q0 = qb * ATT_BR;

    const __nv_bfloat16* base = qkv + static_cast<size_t>(b) * T * ld;
    const __nv_bfloat16* gq = base + h * D;
    const __nv_bfloat16* gk = base + E + h * D;
    const __nv_bfloat16* gv = base + 2 * E + h * D;

    load_tile_async<D, ATT_THREADS>(smem_u32(sQ), gq, ld, q0, ATT_BR, T, tid);
    load_tile_async<D, ATT_THREADS>(smem_u32(sK[0]), gk, ld, 0, ATT_BC, T, tid);
    load_tile_async<D, ATT_THREADS>(smem_u32(sV[0]), gv, ld, 0, ATT_BC, T, tid);
    cp_async_commit();

    float o[NT_O][4];
#pragma unroll
    for (int t = 0; t < NT_O; ++t) { o[t][0] = o[t][1] = o[t][2] = o[t][3] = 0.f; }
    float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
    uint32_t qf[KS][4];

    const int nblocks = qb + 1;
    for (int j = 0; j < nblocks; ++j) {
        const int buf = j & 1;
        if (j + 1 < nblocks) {
            load_tile_async<D, ATT_THREADS>(smem_u32(sK[buf ^ 1]), gk, ld, (j + 1) * ATT_BC, ATT_BC, T, tid);
            load_tile_async<D, ATT_THREADS>(smem_u32(sV[buf ^ 1]), gv, ld, (j + 1) * ATT_BC, ATT_BC, T, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (j == 0) {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = ks * 2 + (lane >> 4);
                ldmatrix_x4(qf[ks], smem_u32(sQ) + tile_off<D>(r, c));
            }
        }
        // ---- S = Q K^T -------------------------------------------------
        float s[8][4];
#pragma unroll
        for (int t = 0; t < 8; ++t) { s[t][0] = s[t][1] = s[t][2] = s[t][3] = 0.f; }
        const uint32_t kbase = smem_u32(sK[buf]);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
            for (int tp = 0; tp < 4; ++tp) {   // pairs of key n-tiles
                uint32_t kf[4];
                const int key = tp * 16 + ((lane >> 4) << 3) + (lane & 7);
                const int c = ks * 2 + ((lane >> 3) & 1);
                ldmatrix_x4(kf, kbase + tile_off<D>(key, c));
                mma_bf16_16816(s[2 * tp], qf[ks], kf[0], kf[1]);
                mma_bf16_16816(s[2 * tp + 1], qf[ks], kf[2], kf[3]);
            }
        }
        // ---- causal mask (only the diagonal block can be partially masked) ----
        if (j == qb) {
            const int r_lo = warp * 16 + g, r_hi = r_lo + 8;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = t * 8 + 2 * tig + e;
                    if (col > r_lo) s[t][e] = -INFINITY;
                    if (col > r_hi) s[t][2 + e] = -INFINITY;
                }
            }
        }
        // ---- online softmax ----------------------------------------------
        float bm_lo = -INFINITY, bm_hi = -INFINITY;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            bm_lo = fmax3(bm_lo, s[t][0], s[t][1]);
            bm_hi = fmax3(bm_hi, s[t][2], s[t][3]);
        }
        bm_lo = fmaxf(bm_lo, __shfl_xor_sync(0xffffffffu, bm_lo, 1));
        bm_lo = fmaxf(bm_lo, __shfl_xor_sync(0xffffffffu, bm_lo, 2));
        bm_hi = fmaxf(bm_hi, __shfl_xor_sync(0xffffffffu, bm_hi, 1));
        bm_hi = fmaxf(bm_hi, __shfl_xor_sync(0xffffffffu, bm_hi, 2));
        const float mn_lo = fmaxf(m_lo, bm_lo), mn_hi = fmaxf(m_hi, bm_hi);
        const float corr_lo = fast_exp2((m_lo - mn_lo) * scale_log2), corr_hi = fast_exp2((m_hi - mn_hi) * scale_log2);
        m_lo = mn_lo; m_hi = mn_hi;
        const float ms_lo = mn_lo * scale_log2, ms_hi = mn_hi * scale_log2;
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
        l_lo = l_lo * corr_lo + ps_lo;
        l_hi = l_hi * corr_hi + ps_hi;
#pragma unroll
        for (int t = 0; t < NT_O; ++t) {
            o[t][0] *= corr_lo; o[t][1] *= corr_lo; o[t][2] *= corr_hi; o[t][3] *= corr_hi;
        }
        // ---- dropout on the probabilities (the row sums above stay undropped) ----
        if (drop.threshold32 != 0) {
            uint32_t x = attn_stream_seed(drop, b * H + h, (q0 >> 4) + warp, j, lane);
#pragma unroll
            for (int t = 0; t < 8; ++t) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    x = x * ATTN_LCG_A + ATTN_LCG_C;
                    if (x < drop.threshold32) s[t][e] = 0.f;
                }
            }
        }
        // ---- O += P V -----------------------------------------------------
        const uint32_t vbase = smem_u32(sV[buf]);
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
                const int key = ks * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
                const int c = np * 2 + (lane >> 4);
                ldmatrix_x4_trans(vf, vbase + tile_off<D>(key, c));
                mma_bf16_16816(o[2 * np], pa, vf[0], vf[1]);
                mma_bf16_16816(o[2 * np + 1], pa, vf[2], vf[3]);
            }
        }
        __syncthreads();   // everyone is done with buf before it is refilled
    }

    // ---- finalize ---------------------------------------------------------
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    const float ks_scale = (drop.threshold32 != 0) ? drop.keep_scale : 1.0f;
    const float inv_lo = ks_scale / l_lo, inv_hi = ks_scale / l_hi;
    const int i_lo = q0 + warp * 16 + g, i_hi = i_lo + 8;
    __nv_bfloat16* ob = out + static_cast<size_t>(b) * T * E + h * D;
#pragma unroll
    for (int t = 0; t < NT_O; ++t) {
        const int col = t * 8 + 2 * tig;
        if (i_lo < T) *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(i_lo) * E + col) = pack_bf16(o[t][0] * inv_lo, o[t][1] * inv_lo);
        if (i_hi < T) *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(i_hi) * E + col) = pack_bf16(o[t][2] * inv_hi, o[t][3] * inv_hi);
    }
    if (lse != nullptr && tig == 0) {
        float* lb = lse + (static_cast<size_t>(b) * H + h) * T;
        if (i_lo < T) lb[i_lo] = m_lo * scale_log2 + log2f(l_lo);
        if (i_hi < T) lb[i_hi] = m_hi * scale_log2 + log2f(l_hi);
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

// ---------------------------------------------------------------------------
// Backward, step 2.  One CTA owns a block of BC keys of one (batch, head) and
// walks the query blocks at or below it.  Each warp takes 16 query rows of the
// 64-row query block, so dQ rows are complete inside a warp (added to the fp32
// dq buffer with vector reductions) while the warp's partial dK / dV stay in
// registers for the whole walk and are combined across the 4 warps at the end.
// The key block is processed in 32-key halves (fewer live registers); a half
// that lies entirely above the warp's rows is skipped, one that straddles the
// diagonal takes the masked path.  P and dS are needed transposed (dV += P^T dO,
// dK += dS^T Q): movmatrix on the packed bf16 accumulator tiles.
// With dropout (keep mask M, keep scale ks): dV = ks * (M.P)^T dO,
// dS = ks * P.(M.dP - delta/ks); the ks factors are applied once at the end.
// ---------------------------------------------------------------------------
template <int D, int BC, bool DROP>
__global__ void __launch_bounds__(ATT_THREADS, (D == 16) ? 4 : 1)
attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                const float* __restrict__ lse, const float* __restrict__ delta, float* __restrict__ dq_acc,
                __nv_bfloat16* __restrict__ dqkv, int T, int H, float scale, float scale_log2, AttnDropKey drop) {
    constexpr int KS = D / 16;
    constexpr int HALVES = BC / 32;
    constexpr int MT = BC / 16;     // m-tiles of dK / dV
    constexpr int NT_D = D / 8;
    constexpr int TILE_Q = ATT_BR * D * 2;
    constexpr int TILE_K = BC * D * 2;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sK = smem;
    uint8_t* sV = sK + TILE_K;
    uint8_t* sQ = sV + TILE_K;             // [2][TILE_Q]
    uint8_t* sdO = sQ + 2 * TILE_Q;        // [2][TILE_Q]
    float* sLse = reinterpret_cast<float*>(sdO + 2 * TILE_Q);   // [2][64]
    float* sDelta = sLse + 2 * ATT_BR;                            // [2][64]
    float* sRed = sDelta + 2 * ATT_BR;                            // [2][BC][D] fp32

    const int E = H * D;
    const int ld = 3 * E;
    const int kb = blockIdx.x;
    const int h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tig = lane & 3;
    const int k0 = kb * BC;
    const int nqb = (T + ATT_BR - 1) / ATT_BR;
    const int qb_first = k0 / ATT_BR;
    const float ks_scale = DROP ? drop.keep_scale : 1.0f;
    const float inv_ks = 1.0f / ks_scale;

    const __nv_bfloat16* base = qkv + static_cast<size_t>(b) * T * ld;
    const __nv_bfloat16* gq = base + h * D;
    const __nv_bfloat16* gk = base + E + h * D;
    const __nv_bfloat16* gv = base + 2 * E + h * D;
    const __nv_bfloat16* gdo = dout + static_cast<size_t>(b) * T * E + h * D;
    const float* glse = lse + (static_cast<size_t>(b) * H + h) * T;
    const float* gdelta = delta + (static_cast<size_t>(b) * H + h) * T;

    auto load_q_block = [&](int qb, int buf) {
        load_tile_async<D, ATT_THREADS>(smem_u32(sQ + buf * TILE_Q), gq, ld, qb * ATT_BR, ATT_BR, T, tid);
        load_tile_async<D, ATT_THREADS>(smem_u32(sdO + buf * TILE_Q), gdo, E, qb * ATT_BR, ATT_BR, T, tid);
        if (tid < ATT_BR) {
            const int r = qb * ATT_BR + tid;
            // rows past the end get lse = +inf so that their probabilities are exactly 0
            sLse[buf * ATT_BR + tid] = (r < T) ? glse[r] : INFINITY;
            sDelta[buf * ATT_BR + tid] = (r < T) ? gdelta[r] * inv_ks : 0.f;
        }
    };

    load_tile_async<D, ATT_THREADS>(smem_u32(sK), gk, ld, k0, BC, T, tid);
    load_tile_async<D, ATT_THREADS>(smem_u32(sV), gv, ld, k0, BC, T, tid);
    load_q_block(qb_first, 0);
    cp_async_commit();

    float dk[MT][NT_D][4], dv[MT][NT_D][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int t = 0; t < NT_D; ++t)
#pragma unroll
            for (int e = 0; e < 4; ++e) { dk[m][t][e] = 0.f; dv[m][t][e] = 0.f; }

    for (int qb = qb_first; qb < nqb; ++qb) {
        const int buf = (qb - qb_first) & 1;
        if (qb + 1 < nqb) {
            load_q_block(qb + 1, buf ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        const uint32_t qbase = smem_u32(sQ + buf * TILE_Q), dobase = smem_u32(sdO + buf * TILE_Q);
        const uint32_t kbase = smem_u32(sK), vbase = smem_u32(sV);
        const int row_min = qb * ATT_BR + warp * 16;     // this warp's first query row
        const int r_lo = warp * 16 + g, r_hi = r_lo + 8;
        const int i_lo = qb * ATT_BR + r_lo, i_hi = i_lo + 8;

        if (k0 <= row_min + 15 && row_min < T) {          // otherwise every key of the block is masked for this warp
            uint32_t qf[KS][4], dof[KS][4];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = ks * 2 + (lane >> 4);
                ldmatrix_x4(qf[ks], qbase + tile_off<D>(r, c));
                ldmatrix_x4(dof[ks], dobase + tile_off<D>(r, c));
            }
            const float lse_lo = sLse[buf * ATT_BR + r_lo], lse_hi = sLse[buf * ATT_BR + r_hi];
            const float dl_lo = sDelta[buf * ATT_BR + r_lo], dl_hi = sDelta[buf * ATT_BR + r_hi];
            float dq[NT_D][4];
#pragma unroll
            for (int t = 0; t < NT_D; ++t) { dq[t][0] = dq[t][1] = dq[t][2] = dq[t][3] = 0.f; }

#pragma unroll
            for (int hf = 0; hf < HALVES; ++hf) {
                const int kh0 = k0 + hf * 32;                 // first key of this half
                if (kh0 > row_min + 15) continue;             // entirely above the diagonal for this warp
                const bool partial = (kh0 + 31) > row_min;    // straddles the diagonal
                // ---- S = Q K^T and dP = dO V^T for 32 keys ------------------
                float s[4][4], dp[4][4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    s[t][0] = s[t][1] = s[t][2] = s[t][3] = 0.f;
                    dp[t][0] = dp[t][1] = dp[t][2] = dp[t][3] = 0.f;
                }
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
                    for (int tp = 0; tp < 2; ++tp) {
                        uint32_t kf[4], vf[4];
                        const int key = hf * 32 + tp * 16 + ((lane >> 4) << 3) + (lane & 7);
                        const int c = ks * 2 + ((lane >> 3) & 1);
                        ldmatrix_x4(kf, kbase + tile_off<D>(key, c));
                        ldmatrix_x4(vf, vbase + tile_off<D>(key, c));
                        mma_bf16_16816(s[2 * tp], qf[ks], kf[0], kf[1]);
                        mma_bf16_16816(s[2 * tp + 1], qf[ks], kf[2], kf[3]);
                        mma_bf16_16816(dp[2 * tp], dof[ks], vf[0], vf[1]);
                        mma_bf16_16816(dp[2 * tp + 1], dof[ks], vf[2], vf[3]);
                    }
                }
                // ---- P = exp2(S*c - lse) ; dS' = P * (M.dP - delta/ks) -------
                uint32_t x = 0;
                if (DROP) {
                    x = attn_stream_seed(drop, b * H + h, row_min >> 4, kh0 >> 6, lane);
                    if (kh0 & 32) x = x * lcg_mul_pow(16) + lcg_add_pow(16);   // second half of the 64-key block
                }
                uint32_t pT[2][4], dsT[2][4];   // [query half][key n-tile], transposed 8x8 blocks
                uint32_t dsA[4][2];             // untransposed dS' for dQ
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    float p[4], ds[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float pv = fast_exp2(fmaf(s[t][e], scale_log2, (e < 2) ? -lse_lo : -lse_hi));
                        if (partial) {
                            const int col = kh0 + t * 8 + 2 * tig + (e & 1);
                            if (col > ((e < 2) ? i_lo : i_hi)) pv = 0.f;
                        }
                        float dpv = dp[t][e];
                        float pd = pv;
                        if (DROP) {
                            x = x * ATTN_LCG_A + ATTN_LCG_C;
                            if (x < drop.threshold32) { dpv = 0.f; pd = 0.f; }
                        }
                        p[e] = pd;
                        ds[e] = pv * (dpv - ((e < 2) ? dl_lo : dl_hi));
                    }
                    const uint32_t p_lo = pack_bf16(p[0], p[1]), p_hi = pack_bf16(p[2], p[3]);
                    const uint32_t d_lo = pack_bf16(ds[0], ds[1]), d_hi = pack_bf16(ds[2], ds[3]);
                    dsA[t][0] = d_lo; dsA[t][1] = d_hi;
                    pT[0][t] = movmatrix_trans(p_lo); pT[1][t] = movmatrix_trans(p_hi);
                    dsT[0][t] = movmatrix_trans(d_lo); dsT[1][t] = movmatrix_trans(d_hi);
                }
                // ---- dQ += dS' K   (k = this half's 32 keys) -----------------
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const uint32_t a[4] = {dsA[2 * ks][0], dsA[2 * ks][1], dsA[2 * ks + 1][0], dsA[2 * ks + 1][1]};
#pragma unroll
                    for (int np = 0; np < NT_D / 2; ++np) {
                        uint32_t kf[4];
                        const int key = hf * 32 + ks * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
                        const int c = np * 2 + (lane >> 4);
                        ldmatrix_x4_trans(kf, kbase + tile_off<D>(key, c));
                        mma_bf16_16816(dq[2 * np], a, kf[0], kf[1]);
                        mma_bf16_16816(dq[2 * np + 1], a, kf[2], kf[3]);
                    }
                }
                // ---- dV += P^T dO ; dK += dS'^T Q   (M = keys, K = this warp's 16 query rows) ----
#pragma unroll
                for (int np = 0; np < NT_D / 2; ++np) {
                    uint32_t dob[4], qb4[4];
                    const int r = warp * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
                    const int c = np * 2 + (lane >> 4);
                    ldmatrix_x4_trans(dob, dobase + tile_off<D>(r, c));
                    ldmatrix_x4_trans(qb4, qbase + tile_off<D>(r, c));
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        const uint32_t pa[4] = {pT[0][2 * m], pT[0][2 * m + 1], pT[1][2 * m], pT[1][2 * m + 1]};
                        const uint32_t da[4] = {dsT[0][2 * m], dsT[0][2 * m + 1], dsT[1][2 * m], dsT[1][2 * m + 1]};
                        mma_bf16_16816(dv[hf * 2 + m][2 * np], pa, dob[0], dob[1]);
                        mma_bf16_16816(dv[hf * 2 + m][2 * np + 1], pa, dob[2], dob[3]);
                        mma_bf16_16816(dk[hf * 2 + m][2 * np], da, qb4[0], qb4[1]);
                        mma_bf16_16816(dk[hf * 2 + m][2 * np + 1], da, qb4[2], qb4[3]);
                    }
                }
            }
            // ---- dQ rows of this warp are complete for this key block ----------
            const float dq_scale = scale * ks_scale;
            float* dqb = dq_acc + static_cast<size_t>(b) * T * E + h * D;
#pragma unroll
            for (int t = 0; t < NT_D; ++t) {
                const int col = t * 8 + 2 * tig;
                if (i_lo < T)
                    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dqb + static_cast<size_t>(i_lo) * E + col),
                                 "f"(dq[t][0] * dq_scale), "f"(dq[t][1] * dq_scale) : "memory");
                if (i_hi < T)
                    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dqb + static_cast<size_t>(i_hi) * E + col),
                                 "f"(dq[t][2] * dq_scale), "f"(dq[t][3] * dq_scale) : "memory");
            }
        }
        __syncthreads();
    }

    // ---- combine the 4 warps' dK / dV partials and store bf16 -----------------
    for (int i = tid; i < 2 * BC * D; i += ATT_THREADS) sRed[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int t = 0; t < NT_D; ++t)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int key = m * 16 + g + ((e >> 1) << 3);
                const int col = t * 8 + 2 * tig + (e & 1);
                atomicAdd(&sRed[key * D + col], dk[m][t][e]);
                atomicAdd(&sRed[BC * D + key * D + col], dv[m][t][e]);
            }
    __syncthreads();
    const float dk_scale = scale * ks_scale;
    __nv_bfloat16* dkb = dqkv + static_cast<size_t>(b) * T * ld + E + h * D;
    __nv_bfloat16* dvb = dqkv + static_cast<size_t>(b) * T * ld + 2 * E + h * D;
    for (int i = tid; i < BC * D / 2; i += ATT_THREADS) {
        const int key = (2 * i) / D, col = (2 * i) % D;
        if (k0 + key < T) {
            *reinterpret_cast<uint32_t*>(dkb + static_cast<size_t>(k0 + key) * ld + col) =
                pack_bf16(sRed[key * D + col] * dk_scale, sRed[key * D + col + 1] * dk_scale);
            *reinterpret_cast<uint32_t*>(dvb + static_cast<size_t>(k0 + key) * ld + col) =
                pack_bf16(sRed[BC * D + key * D + col] * ks_scale, sRed[BC * D + key * D + col + 1] * ks_scale);
        }
    }
}

// dq (fp32 accumulation buffer, [rows, E]) -> bf16 into the q third of dqkv ([rows, 3E]); re-zeroes the buffer.
__global__ void __launch_bounds__(256)
attn_dq_store_kernel(float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dqkv, size_t rows, int E) {
    const size_t n4 = rows * (E / 4);
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t row = i / (E / 4);
        const int c = static_cast<int>(i % (E / 4)) * 4;
        float4* src = reinterpret_cast<float4*>(dq_acc + row * E + c);
        const float4 v = *src;
        *src = make_float4(0.f, 0.f, 0.f, 0.f);
        uint2 o;
        o.x = pack_bf16(v.x, v.y); o.y = pack_bf16(v.z, v.w);
        *reinterpret_cast<uint2*>(dqkv + row * 3 * E + c) = o;
    }
}

// ---------------------------------------------------------------------------
// Host launchers
// ---------------------------------------------------------------------------
static const float kLog2e = 1.4426950408889634f;

int attention_fwd(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse, int B, int T, int H, int D, float scale,
                  const DropoutParams& drop, uint32_t layer, cudaStream_t s) {
    if (B * T == 0) return 0;
    CB200_REQUIRE(T < (1 << 17), "sequences of 2^17 tokens or more are not supported by the attention dropout stream");
    dim3 grid((T + ATT_BR - 1) / ATT_BR, H, B);
    const float c = scale * kLog2e;
    const AttnDropKey key = make_attn_drop_key(drop, layer);
    switch (D) {
        case 16: attn_fwd_kernel<16><<<grid, ATT_THREADS, 0, s>>>(qkv, out, lse, T, H, c, key); break;
        case 32: attn_fwd_kernel<32><<<grid, ATT_THREADS, 0, s>>>(qkv, out, lse, T, H, c, key); break;
        case 64: attn_fwd_kernel<64><<<grid, ATT_THREADS, 0, s>>>(qkv, out, lse, T, H, c, key); break;
        default: set_error("attention head size %d is not supported (16, 32 or 64)", D); return -1;
    }
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

template <int D, int BC, bool DROP>
static int launch_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* dout, const float* lse, const float* delta,
                      float* dq_acc, __nv_bfloat16* dqkv, int B, int T, int H, float scale, const AttnDropKey& key,
                      cudaStream_t s) {
    constexpr size_t smem = 2 * BC * D * 2 + 4 * ATT_BR * D * 2 + 4 * ATT_BR * sizeof(float) + 2 * BC * D * sizeof(float);
    auto kernel = attn_bwd_kernel<D, BC, DROP>;
    static bool configured = false;
    if (!configured) {
        CB200_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid((T + BC - 1) / BC, H, B);
    kernel<<<grid, ATT_THREADS, smem, s>>>(qkv, dout, lse, delta, dq_acc, dqkv, T, H, scale, scale * kLog2e, key);
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

template <int D, int BC>
static int launch_bwd_drop(bool dropping, const __nv_bfloat16* qkv, const __nv_bfloat16* dout, const float* lse,
                           const float* delta, float* dq_acc, __nv_bfloat16* dqkv, int B, int T, int H, float scale,
                           const AttnDropKey& key, cudaStream_t s) {
    return dropping ? launch_bwd<D, BC, true>(qkv, dout, lse, delta, dq_acc, dqkv, B, T, H, scale, key, s)
                    : launch_bwd<D, BC, false>(qkv, dout, lse, delta, dq_acc, dqkv, B, T, H, scale, key, s);
}

int attention_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* out, const __nv_bfloat16* dout, const float* lse,
                  float* delta, float* dq_acc, __nv_bfloat16* dqkv, int B, int T, int H, int D, float scale,
                  const DropoutParams& drop, uint32_t layer, cudaStream_t s) {
    if (B * T == 0) return 0;
    const int rows = B * T;
    const int E = H * D;
    const AttnDropKey key = make_attn_drop_key(drop, layer);
    const bool dropping = key.threshold32 != 0;
    int rc = 0;
    switch (D) {
        case 16:
            attn_bwd_delta_kernel<16><<<(rows + 7) / 8, 256, 0, s>>>(dout, out, delta, rows, T, H);
            rc = launch_bwd_drop<16, 64>(dropping, qkv, dout, lse, delta, dq_acc, dqkv, B, T, H, scale, key, s);
            break;
        case 32:
            attn_bwd_delta_kernel<32><<<(rows + 7) / 8, 256, 0, s>>>(dout, out, delta, rows, T, H);
            rc = launch_bwd_drop<32, 32>(dropping, qkv, dout, lse, delta, dq_acc, dqkv, B, T, H, scale, key, s);
            break;
        case 64:
            attn_bwd_delta_kernel<64><<<(rows + 7) / 8, 256, 0, s>>>(dout, out, delta, rows, T, H);
            rc = launch_bwd_drop<64, 32>(dropping, qkv, dout, lse, delta, dq_acc, dqkv, B, T, H, scale, key, s);
            break;
        default: set_error("attention head size %d is not supported (16, 32 or 64)", D); return -1;
    }
    if (rc) return rc;
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    size_t n4 = static_cast<size_t>(rows) * (E / 4);
    size_t blocks = (n4 + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    attn_dq_store_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(dq_acc, dqkv, rows, E);
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

// Debug/parity helper: materialise the attention-probability keep mask
// ([B, H, T, T] bytes, 1 = kept) that the kernels above apply.  One thread per
// (16-row group, 64-key block, lane) replays that lane's LCG stream.
__global__ void attn_mask_export_kernel(uint8_t* __restrict__ mask, int T, int H, AttnDropKey drop) {
    const int bh = blockIdx.y;
    const int n16 = (T + 15) / 16, n64 = (T + 63) / 64;
    const int total = n16 * n64 * 32;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int lane = idx & 31, jb = (idx >> 5) % n64, i16 = (idx >> 5) / n64;
        const int g = lane >> 2, tig = lane & 3;
        uint32_t x = attn_stream_seed(drop, bh, i16, jb, lane);
        for (int t = 0; t < 8; ++t)
            for (int e = 0; e < 4; ++e) {
                x = x * ATTN_LCG_A + ATTN_LCG_C;
                const int i = i16 * 16 + g + 8 * (e >> 1), j = jb * 64 + 8 * t + 2 * tig + (e & 1);
                if (i < T && j < T)
                    mask[(static_cast<size_t>(bh) * T + i) * T + j] = (drop.threshold32 == 0 || x >= drop.threshold32) ? 1 : 0;
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
