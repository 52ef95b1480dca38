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
// At the default d_h = 16 one 64x64 score block is 16 m16n8k16 MMAs but 4096
// exponentials plus the softmax / dropout arithmetic around them: ncu shows
// the kernels issue-bound on the FMA/ALU/XU pipes with the tensor pipe ~25 %
// busy, so the design goal is instructions per score element, not tensor
// throughput.  Scores are therefore produced with warp-level mma.sync straight
// into registers in the layout the softmax consumes (no TMEM round trip), a warp
// owns whole query rows (no cross-warp softmax traffic), and the dropout mask
// costs 3 instructions per element (see common.cuh).  See DESIGN.md "Attention".
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
    AttnStream seed_base{0u, 1u};
    if (DROP) seed_base = attn_stream_base(drop, b * H + h, lane);

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
                if (DROP) {
                    uint32_t x = attn_stream_seed(seed_base, row0 >> 4, j);
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            x *= ATTN_MCG_A;
                            if (x < drop.threshold32) s[t][e] = 0.f;
                        }
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

// ---------------------------------------------------------------------------
// Backward, step 2.  One CTA owns a block of BC keys of one (batch, head) and
// walks the query blocks (64*MW rows) at or below it.  Each warp takes MW
// groups of 16 query rows, so dQ rows are complete inside a warp (added to the
// fp32 dq buffer with vector reductions) while the warp's partial dK / dV stay
// in registers for the whole walk and are combined across the 4 warps at the
// end.  The key block is processed in 32-key halves (fewer live registers); a
// half that lies entirely above the group's rows is skipped, one that straddles
// the diagonal takes the masked path.  P and dS are needed transposed
// (dV += P^T dO, dK += dS^T Q): movmatrix on the packed bf16 accumulator tiles.
// With dropout (keep mask M, keep scale ks): dV = ks * (M.P)^T dO,
// dS = ks * P.(M.dP - delta/ks); the ks factors are applied once at the end.
// ---------------------------------------------------------------------------
template <int D, int BC, int MW, bool DROP, int MINB>
__global__ void __launch_bounds__(ATT_THREADS, MINB)
attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                const float* __restrict__ lse, const float* __restrict__ delta, float* __restrict__ dq_acc,
                __nv_bfloat16* __restrict__ dqkv, int T, int H, float scale, float scale_log2, AttnDropKey drop) {
    constexpr int BR = 64 * MW;
    constexpr int KS = D / 16;
    constexpr int HALVES = BC / 32;
    constexpr int MT = BC / 16;     // m-tiles of dK / dV
    constexpr int NT_D = D / 8;
    constexpr int TILE_Q = BR * D * 2;
    constexpr int TILE_K = BC * D * 2;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sK = smem;
    uint8_t* sV = sK + TILE_K;
    uint8_t* sQ = sV + TILE_K;             // [2][TILE_Q]
    uint8_t* sdO = sQ + 2 * TILE_Q;        // [2][TILE_Q]
    float* sLse = reinterpret_cast<float*>(sdO + 2 * TILE_Q);   // [2][BR]
    float* sDelta = sLse + 2 * BR;                                // [2][BR]
    float* sRed = sDelta + 2 * BR;                                // [2][BC][D] fp32

    const int E = H * D;
    const int ld = 3 * E;
    const int kb = blockIdx.x;
    const int h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int g = lane >> 2, tig = lane & 3;
    const int k0 = kb * BC;
    const int nqb = (T + BR - 1) / BR;
    const int qb_first = k0 / BR;
    const float ks_scale = DROP ? drop.keep_scale : 1.0f;
    const float inv_ks = 1.0f / ks_scale;

    const __nv_bfloat16* base = qkv + static_cast<size_t>(b) * T * ld;
    const __nv_bfloat16* gq = base + h * D;
    const __nv_bfloat16* gk = base + E + h * D;
    const __nv_bfloat16* gv = base + 2 * E + h * D;
    const __nv_bfloat16* gdo = dout + static_cast<size_t>(b) * T * E + h * D;
    const float* glse = lse + (static_cast<size_t>(b) * H + h) * T;
    const float* gdelta = delta + (static_cast<size_t>(b) * H + h) * T;

    TileCopy<D, BR> q_copy;
    q_copy.init(tid);
    auto load_q_block = [&](int qb, int buf) {
        q_copy.issue(smem_u32(sQ + buf * TILE_Q), gq, ld, qb * BR, T);
        q_copy.issue(smem_u32(sdO + buf * TILE_Q), gdo, E, qb * BR, T);
        if (tid < BR) {
            const int r = qb * BR + tid;
            // rows past the end get lse = +inf so that their probabilities are exactly 0
            sLse[buf * BR + tid] = (r < T) ? glse[r] : INFINITY;
            sDelta[buf * BR + tid] = (r < T) ? gdelta[r] * inv_ks : 0.f;
        }
    };
    {
        TileCopy<D, BC> k_copy;
        k_copy.init(tid);
        k_copy.issue(smem_u32(sK), gk, ld, k0, T);
        k_copy.issue(smem_u32(sV), gv, ld, k0, T);
    }
    load_q_block(qb_first, 0);
    cp_async_commit();

    float dk[MT][NT_D][4], dv[MT][NT_D][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int t = 0; t < NT_D; ++t)
#pragma unroll
            for (int e = 0; e < 4; ++e) { dk[m][t][e] = 0.f; dv[m][t][e] = 0.f; }
    AttnStream seed_base{0u, 1u};
    if (DROP) seed_base = attn_stream_base(drop, b * H + h, lane);

    const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_chunk = lane >> 4;               // A fragments (Q, dO)
    const int k_row = ((lane >> 4) << 3) + (lane & 7), k_chunk = (lane >> 3) & 1;            // B fragments, k = d
    const int t_row = (((lane >> 3) & 1) << 3) + (lane & 7), t_chunk = lane >> 4;            // B fragments, transposed
    const uint32_t kbase = smem_u32(sK), vbase = smem_u32(sV);
    const float dq_scale = scale * ks_scale;
    float* dqb = dq_acc + static_cast<size_t>(b) * T * E + h * D;

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

#pragma unroll
        for (int mt = 0; mt < MW; ++mt) {
            const int lrow0 = (warp * MW + mt) * 16;           // first row of the group inside the query block
            const int row0 = qb * BR + lrow0;
            if (k0 <= row0 + 15 && row0 < T) {                 // otherwise every key of the block is masked for the group
                uint32_t qf[KS][4], dof[KS][4];
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    ldmatrix_x4(qf[ks], qbase + tile_off<D>(lrow0 + a_row, ks * 2 + a_chunk));
                    ldmatrix_x4(dof[ks], dobase + tile_off<D>(lrow0 + a_row, ks * 2 + a_chunk));
                }
                const float lse_lo = sLse[buf * BR + lrow0 + g], lse_hi = sLse[buf * BR + lrow0 + g + 8];
                const float dl_lo = sDelta[buf * BR + lrow0 + g], dl_hi = sDelta[buf * BR + lrow0 + g + 8];
                const int i_lo = row0 + g, i_hi = i_lo + 8;
                float dq[NT_D][4];
#pragma unroll
                for (int t = 0; t < NT_D; ++t) { dq[t][0] = dq[t][1] = dq[t][2] = dq[t][3] = 0.f; }
                uint32_t x = 0;   // dropout stream state, runs across the halves of a 64-key block

#pragma unroll
                for (int hf = 0; hf < HALVES; ++hf) {
                    const int kh0 = k0 + hf * 32;                 // first key of this half
                    if (kh0 <= row0 + 15) {                       // otherwise entirely above the diagonal
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
                                ldmatrix_x4(kf, kbase + tile_off<D>(hf * 32 + tp * 16 + k_row, ks * 2 + k_chunk));
                                ldmatrix_x4(vf, vbase + tile_off<D>(hf * 32 + tp * 16 + k_row, ks * 2 + k_chunk));
                                mma_bf16_16816(s[2 * tp], qf[ks], kf[0], kf[1]);
                                mma_bf16_16816(s[2 * tp + 1], qf[ks], kf[2], kf[3]);
                                mma_bf16_16816(dp[2 * tp], dof[ks], vf[0], vf[1]);
                                mma_bf16_16816(dp[2 * tp + 1], dof[ks], vf[2], vf[3]);
                            }
                        }
                        // ---- P = exp2(S*c - lse) (+ causal mask on the diagonal) -----
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            s[t][0] = fast_exp2(fmaf(s[t][0], scale_log2, -lse_lo));
                            s[t][1] = fast_exp2(fmaf(s[t][1], scale_log2, -lse_lo));
                            s[t][2] = fast_exp2(fmaf(s[t][2], scale_log2, -lse_hi));
                            s[t][3] = fast_exp2(fmaf(s[t][3], scale_log2, -lse_hi));
                        }
                        if (kh0 + 31 > row0) {
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const int col = kh0 + t * 8 + 2 * tig + e;
                                    if (col > i_lo) s[t][e] = 0.f;
                                    if (col > i_hi) s[t][2 + e] = 0.f;
                                }
                            }
                        }
                        // ---- dropout, dS' = P * (M.dP - delta/ks), pack, transpose ----
                        if (DROP && (hf == 0 || HALVES == 1)) {
                            x = attn_stream_seed(seed_base, row0 >> 4, kh0 >> 6);
                            if (kh0 & 32) x *= mcg_mul_pow(16);   // second half of a 64-key dropout block
                        }
                        uint32_t pT[2][4], dsT[2][4];   // [query half][key n-tile], transposed 8x8 blocks
                        uint32_t dsA[4][2];             // untransposed dS' for dQ
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            float pd[4], ds[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float dpv = dp[t][e];
                                pd[e] = s[t][e];
                                if (DROP) {
                                    x *= ATTN_MCG_A;
                                    const bool dropped = x < drop.threshold32;
                                    dpv = dropped ? 0.f : dpv;
                                    pd[e] = dropped ? 0.f : pd[e];
                                }
                                ds[e] = s[t][e] * (dpv - ((e < 2) ? dl_lo : dl_hi));
                            }
                            const uint32_t p_lo = pack_bf16(pd[0], pd[1]), p_hi = pack_bf16(pd[2], pd[3]);
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
                                ldmatrix_x4_trans(kf, kbase + tile_off<D>(hf * 32 + ks * 16 + t_row, np * 2 + t_chunk));
                                mma_bf16_16816(dq[2 * np], a, kf[0], kf[1]);
                                mma_bf16_16816(dq[2 * np + 1], a, kf[2], kf[3]);
                            }
                        }
                        // ---- dV += P^T dO ; dK += dS'^T Q   (M = keys, K = the group's 16 query rows) ----
#pragma unroll
                        for (int np = 0; np < NT_D / 2; ++np) {
                            uint32_t dob[4], qb4[4];
                            ldmatrix_x4_trans(dob, dobase + tile_off<D>(lrow0 + t_row, np * 2 + t_chunk));
                            ldmatrix_x4_trans(qb4, qbase + tile_off<D>(lrow0 + t_row, np * 2 + t_chunk));
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
                }
                // ---- dQ rows of this group are complete for this key block ----------
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
static int g_attention_bwd_impl = 0;
void attention_set_bwd_impl(int impl) { g_attention_bwd_impl = impl; }

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
    const float c = scale * kLog2e;
    const AttnDropKey key = make_attn_drop_key(drop, layer);
    const bool dropping = key.threshold32 != 0;
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

template <int D, int BC, int MW, bool DROP, int MINB>
static int launch_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* dout, const float* lse, const float* delta,
                      float* dq_acc, __nv_bfloat16* dqkv, int B, int T, int H, float scale, const AttnDropKey& key,
                      cudaStream_t s) {
    constexpr int BR = 64 * MW;
    constexpr size_t smem = 2 * BC * D * 2 + 4 * BR * D * 2 + 4 * BR * sizeof(float) + 2 * BC * D * sizeof(float);
    auto kernel = attn_bwd_kernel<D, BC, MW, DROP, MINB>;
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

template <int D, int BC, int MW, int MINB>
static int launch_bwd_drop(bool dropping, const __nv_bfloat16* qkv, const __nv_bfloat16* dout, const float* lse,
                           const float* delta, float* dq_acc, __nv_bfloat16* dqkv, int B, int T, int H, float scale,
                           const AttnDropKey& key, cudaStream_t s) {
    return dropping ? launch_bwd<D, BC, MW, true, MINB>(qkv, dout, lse, delta, dq_acc, dqkv, B, T, H, scale, key, s)
                    : launch_bwd<D, BC, MW, false, MINB>(qkv, dout, lse, delta, dq_acc, dqkv, B, T, H, scale, key, s);
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
    const bool use_tc = g_attention_bwd_impl == 0;   // 0: tcgen05 / TMEM (default), 1: warp-level mma.sync
    switch (D) {
        case 16: attn_bwd_delta_kernel<16><<<(rows + 7) / 8, 256, 0, s>>>(dout, out, delta, rows, T, H); break;
        case 32: attn_bwd_delta_kernel<32><<<(rows + 7) / 8, 256, 0, s>>>(dout, out, delta, rows, T, H); break;
        case 64: attn_bwd_delta_kernel<64><<<(rows + 7) / 8, 256, 0, s>>>(dout, out, delta, rows, T, H); break;
        default: set_error("attention head size %d is not supported (16, 32 or 64)", D); return -1;
    }
    if (use_tc) {
        rc = attention_bwd_tc_main(qkv, dout, lse, delta, dq_acc, dqkv, B, T, H, D, scale, key, s);
    } else if (D == 16) {
        rc = launch_bwd_drop<16, 64, 2, 4>(dropping, qkv, dout, lse, delta, dq_acc, dqkv, B, T, H, scale, key, s);
    } else if (D == 32) {
        rc = launch_bwd_drop<32, 32, 1, 1>(dropping, qkv, dout, lse, delta, dq_acc, dqkv, B, T, H, scale, key, s);
    } else {
        rc = launch_bwd_drop<64, 32, 1, 1>(dropping, qkv, dout, lse, delta, dq_acc, dqkv, B, T, H, scale, key, s);
    }
    if (rc) return rc;
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    size_t n8 = static_cast<size_t>(rows) * (E / 8);
    size_t blocks = (n8 + 4 * 256 - 1) / (4 * 256);
    if (blocks > 8192) blocks = 8192;
    attn_dq_store_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(dq_acc, dqkv, rows, E);
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

// Debug/parity helper: materialise the attention-probability keep mask
// ([B, H, T, T] bytes, 1 = kept) that the kernels above apply.  One thread per
// (16-row group, 64-key block, lane) replays that lane's stream.
__global__ void attn_mask_export_kernel(uint8_t* __restrict__ mask, int T, int H, AttnDropKey drop) {
    const int bh = blockIdx.y;
    const int n16 = (T + 15) / 16, n64 = (T + 63) / 64;
    const int total = n16 * n64 * 32;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int lane = idx & 31, jb = (idx >> 5) % n64, i16 = (idx >> 5) / n64;
        const int g = lane >> 2, tig = lane & 3;
        uint32_t x = attn_stream_seed(attn_stream_base(drop, bh, lane), i16, jb);
        for (int t = 0; t < 8; ++t)
            for (int e = 0; e < 4; ++e) {
                x *= ATTN_MCG_A;
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
