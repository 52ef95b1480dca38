// Autoregressive decoding kernels: single-query attention over an in-place
// KV cache and the fused temperature / softmax / Philox multinomial sampler.
//
// Reference: the model's `past=` path (composer/models/transformer.py:423-437,
// :735-770) re-allocates and copies the whole cache every step (tf.concat); here
// the cache is a preallocated [L][2][B][H][T_max][D] bf16 tensor appended in
// place.  Sampling replaces cli.py:670-673 (logits / temperature,
// tf.random.categorical, last position).
#include "decode.h"
#include "decode_common.cuh"
#include "mma_sync.cuh"

namespace cb200 {

// One CTA (4 warps) per (sequence, head).  The new token's k, v (from the
// c_attn output row) are appended at position `pos`, then the query attends
// over positions [0, pos].  Keys are streamed as 16-byte chunks: a head row of
// D bf16 is CH = D/8 chunks, consecutive lanes take consecutive chunks, so every
// warp load instruction covers 512 contiguous bytes of the cache.  The CH
// lanes that share a key combine their partial dot products with shuffles; the
// softmax is kept online per lane group and merged with warp shuffles, then
// across the 4 warps through shared memory.
template <int D>
__global__ void __launch_bounds__(128)
decode_attn_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ kcache,
                   __nv_bfloat16* __restrict__ vcache, __nv_bfloat16* __restrict__ out, const int* __restrict__ pos_ptr,
                   int H, int t_max, float scale_log2) {
    constexpr int CH = D / 8;            // 16-byte chunks per key
    constexpr int KPW = 32 / CH;         // keys per warp iteration
    const int h = blockIdx.x, b = blockIdx.y;
    const int E = H * D;
    const int pos = *pos_ptr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int part = lane % CH;          // which 8-wide slice of the head this lane owns
    const size_t head_base = (static_cast<size_t>(b) * H + h) * t_max * D;
    const __nv_bfloat16* row = qkv + static_cast<size_t>(b) * 3 * E + h * D;

    // append k, v of the new token (each lane group writes once; warp 0 only)
    if (warp == 0 && lane < 2 * CH) {
        const int which = lane / CH;     // 0 = k, 1 = v
        const uint4 val = *reinterpret_cast<const uint4*>(row + (1 + which) * E + part * 8);
        __nv_bfloat16* dst = (which == 0 ? kcache : vcache) + head_base + static_cast<size_t>(pos) * D + part * 8;
        *reinterpret_cast<uint4*>(dst) = val;
    }
    float q[8];
    {
        const uint4 qv = *reinterpret_cast<const uint4*>(row + part * 8);
        const float2 a0 = unpack_bf16(qv.x), a1 = unpack_bf16(qv.y), a2 = unpack_bf16(qv.z), a3 = unpack_bf16(qv.w);
        q[0] = a0.x; q[1] = a0.y; q[2] = a1.x; q[3] = a1.y; q[4] = a2.x; q[5] = a2.y; q[6] = a3.x; q[7] = a3.y;
    }
    __syncthreads();   // the appended row is visible to the whole CTA (same-CTA global write + barrier)

    float m = -INFINITY, l = 0.f, acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int n_keys = pos + 1;
    const __nv_bfloat16* kb = kcache + head_base;
    const __nv_bfloat16* vb = vcache + head_base;
    constexpr int UNROLL = 4;            // key groups in flight per warp: 2 * UNROLL 16-byte loads per lane
    for (int key0 = warp * KPW; key0 < n_keys; key0 += 4 * KPW * UNROLL) {
        uint4 kv[UNROLL], vv[UNROLL];
        bool valid[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int key = key0 + u * 4 * KPW + lane / CH;
            valid[u] = key < n_keys;
            kv[u] = make_uint4(0, 0, 0, 0); vv[u] = make_uint4(0, 0, 0, 0);
            if (valid[u]) {
                // the row appended above was written through the normal path: read it coherently
                if (key == pos) {
                    kv[u] = *reinterpret_cast<const uint4*>(kb + static_cast<size_t>(key) * D + part * 8);
                    vv[u] = *reinterpret_cast<const uint4*>(vb + static_cast<size_t>(key) * D + part * 8);
                } else {
                    kv[u] = ld_nc_v4(kb + static_cast<size_t>(key) * D + part * 8);
                    vv[u] = ld_nc_v4(vb + static_cast<size_t>(key) * D + part * 8);
                }
            }
        }
        float sc[UNROLL];
        float mn = m;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            float s = dot8(kv[u], q);
#pragma unroll
            for (int o = 1; o < CH; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            sc[u] = valid[u] ? s * scale_log2 : -INFINITY;
            mn = fmaxf(mn, sc[u]);
        }
        if (mn != -INFINITY) {
            const float corr = fast_exp2(m - mn);
            l *= corr;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] *= corr;
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const float p = fast_exp2(sc[u] - mn);
                l += p;
                const float2 v0 = unpack_bf16(vv[u].x), v1 = unpack_bf16(vv[u].y), v2 = unpack_bf16(vv[u].z), v3 = unpack_bf16(vv[u].w);
                acc[0] += p * v0.x; acc[1] += p * v0.y; acc[2] += p * v1.x; acc[3] += p * v1.y;
                acc[4] += p * v2.x; acc[5] += p * v2.y; acc[6] += p * v3.x; acc[7] += p * v3.y;
            }
            m = mn;
        }
    }
    // merge lanes that own the same slice (`part`) of the head: xor over CH, 2CH, ...
#pragma unroll
    for (int o = CH; o < 32; o <<= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
        const float l2 = __shfl_xor_sync(0xffffffffu, l, o);
        const float mn = fmaxf(m, m2);
        const float c1 = (m == -INFINITY) ? 0.f : fast_exp2(m - mn);
        const float c2 = (m2 == -INFINITY) ? 0.f : fast_exp2(m2 - mn);
        l = l * c1 + l2 * c2;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float a2 = __shfl_xor_sync(0xffffffffu, acc[e], o);
            acc[e] = acc[e] * c1 + a2 * c2;
        }
        m = mn;
    }
    __shared__ float sm[4], sl[4], sacc[4][D];
    if (lane < CH) {
        if (lane == 0) { sm[warp] = m; sl[warp] = l; }
#pragma unroll
        for (int e = 0; e < 8; ++e) sacc[warp][part * 8 + e] = acc[e];
    }
    __syncthreads();
    if (tid < D) {
        float mm = fmaxf(fmaxf(sm[0], sm[1]), fmaxf(sm[2], sm[3]));
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const float c = (sm[w] == -INFINITY) ? 0.f : fast_exp2(sm[w] - mm);
            num += sacc[w][tid] * c;
            den += sl[w] * c;
        }
        out[static_cast<size_t>(b) * E + h * D + tid] = __float2bfloat16_rn(num / den);
    }
}

int decode_attention(const __nv_bfloat16* qkv, __nv_bfloat16* kcache, __nv_bfloat16* vcache, __nv_bfloat16* out,
                     const int* pos_ptr, int B, int H, int D, int t_max, float scale, cudaStream_t s) {
    if (B == 0) return 0;
    dim3 grid(H, B);
    const float c = scale * 1.4426950408889634f;
    switch (D) {
        case 16: decode_attn_kernel<16><<<grid, 128, 0, s>>>(qkv, kcache, vcache, out, pos_ptr, H, t_max, c); break;
        case 32: decode_attn_kernel<32><<<grid, 128, 0, s>>>(qkv, kcache, vcache, out, pos_ptr, H, t_max, c); break;
        case 64: decode_attn_kernel<64><<<grid, 128, 0, s>>>(qkv, kcache, vcache, out, pos_ptr, H, t_max, c); break;
        default: set_error("attention head size %d is not supported (16, 32 or 64)", D); return -1;
    }
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

// Token + positional embedding for one decode step: row b gets wte[cur[b]] + wpe[*pos].
__global__ void __launch_bounds__(256)
decode_embed_kernel(const int32_t* __restrict__ cur, const float* __restrict__ wte, const float* __restrict__ wpe,
                    __nv_bfloat16* __restrict__ out, const int* __restrict__ pos_ptr, int B, int E, int vocab) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= B) return;
    const int lane = threadIdx.x & 31;
    const int pos = *pos_ptr;
    int id = cur[row];
    id = min(max(id, 0), vocab - 1);
    const float* te = wte + static_cast<size_t>(id) * E;
    const float* pe = wpe + static_cast<size_t>(pos) * E;
    for (int c4 = lane; c4 < E / 4; c4 += 32) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(te + c4 * 4));
        const float4 p = __ldg(reinterpret_cast<const float4*>(pe + c4 * 4));
        uint2 o;
        o.x = pack_bf16(a.x + p.x, a.y + p.y); o.y = pack_bf16(a.z + p.z, a.w + p.w);
        *reinterpret_cast<uint2*>(out + static_cast<size_t>(row) * E + c4 * 4) = o;
    }
}

// ---- batched prefill: k, v of T prompt tokens from the c_attn output into the per-step cache layout ----
// qkv: [B * T, 3E] (q | k | v thirds, head h at columns h D); kcache / vcache: [B, H, t_max, D] of one layer
// (transformer.py:419-426: split_heads of key and value, i.e. the reference's `present` for positions 0 .. T-1).
// One thread per 16-byte piece.
__global__ void __launch_bounds__(256)
kv_export_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ kcache,
                 __nv_bfloat16* __restrict__ vcache, int B, int T, int H, int D, int t_max) {
    const int pieces = D / 8, E = H * D;
    const size_t n = static_cast<size_t>(B) * T * H * pieces * 2;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        size_t r = i;
        const int c = static_cast<int>(r % pieces); r /= pieces;
        const int h = static_cast<int>(r % H); r /= H;
        const int kv = static_cast<int>(r & 1); r >>= 1;
        const int t = static_cast<int>(r % T);
        const int b = static_cast<int>(r / T);
        const uint4 val = *reinterpret_cast<const uint4*>(qkv + (static_cast<size_t>(b) * T + t) * 3 * E + (1 + kv) * E + h * D + c * 8);
        __nv_bfloat16* dst = (kv ? vcache : kcache) + ((static_cast<size_t>(b) * H + h) * t_max + t) * D + c * 8;
        *reinterpret_cast<uint4*>(dst) = val;
    }
}

int kv_export(const __nv_bfloat16* qkv, __nv_bfloat16* kcache, __nv_bfloat16* vcache, int B, int T, int H, int D,
              int t_max, cudaStream_t s) {
    const size_t n = static_cast<size_t>(B) * T * H * (D / 8) * 2;
    const int blocks = static_cast<int>(n / 256 + 1 < 148 * 8 ? n / 256 + 1 : 148 * 8);
    kv_export_kernel<<<blocks, 256, 0, s>>>(qkv, kcache, vcache, B, T, H, D, t_max);
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

int decode_embed(const int32_t* cur, const float* wte, const float* wpe, __nv_bfloat16* out, const int* pos_ptr, int B,
                 int E, int vocab, cudaStream_t s) {
    if (B == 0) return 0;
    decode_embed_kernel<<<(B + 7) / 8, 256, 0, s>>>(cur, wte, wpe, out, pos_ptr, B, E, vocab);
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

// Advances the device-side position / step counters once per launch (last block to finish).
__device__ __forceinline__ void advance_counters(int* pos_ptr, int* step_ptr, int step) {
    __shared__ int block_done;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int ticket = atomicAdd(reinterpret_cast<unsigned int*>(pos_ptr + 1), 1u);
        block_done = (ticket == gridDim.x * gridDim.y - 1);
    }
    __syncthreads();
    if (block_done && threadIdx.x == 0) {
        pos_ptr[1] = 0;          // reset the ticket
        *pos_ptr = *pos_ptr + 1;
        *step_ptr = step + 1;
    }
}

// Sampling from materialised logits (one warp per sequence).  When `forced` is non-null and
// forced[b*forced_ld + step] >= 0 that id is emitted instead (prompt teacher-forcing).
__global__ void __launch_bounds__(128)
sample_kernel(const float* __restrict__ logits, int ld, int V, float inv_temperature, int greedy, uint32_t seed_lo,
              uint32_t seed_hi, int seq_base, int32_t* __restrict__ out_ids, int out_ld, int32_t* __restrict__ cur,
              const int32_t* __restrict__ forced, int forced_ld, int* __restrict__ pos_ptr, int* __restrict__ step_ptr,
              float* __restrict__ u_out, int B) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + warp;
    const int step = *step_ptr;
    if (b < B) {
        float u;
        int chosen = sample_row(logits + static_cast<size_t>(b) * ld, V, inv_temperature, greedy, seed_lo, seed_hi,
                                static_cast<uint32_t>(seq_base + b), static_cast<uint32_t>(step), lane, &u);
        if (forced != nullptr) {
            const int f = forced[static_cast<size_t>(b) * forced_ld + step];
            if (f >= 0) chosen = f;
        }
        if (lane == 0) {
            out_ids[static_cast<size_t>(b) * out_ld + step] = chosen;
            cur[b] = chosen;
            if (u_out != nullptr) u_out[static_cast<size_t>(b) * out_ld + step] = u;
        }
    }
    advance_counters(pos_ptr, step_ptr, step);
}

// ln_f + tied logits + sampling for one decode step, one CTA (4 warps) per sequence: the final hidden
// row is normalised in shared memory, every warp takes vocabulary rows v = warp, warp + 4, ... (a lane
// reads 16 contiguous bytes of wte[v], so a warp load covers 512 contiguous bytes), the 390 logits stay
// in shared memory and warp 0 draws the token.  Replaces LayerNorm + logits GEMM + sampler launches.
// Reference: transformer.py:811, :818 (tied logits) and cli.py:670-673.
template <int VPL>   // 8-element vectors per lane over E: E = 256 * VPL
__global__ void __launch_bounds__(128)
logits_sample_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, const __nv_bfloat16* __restrict__ wte, int V,
                     float inv_temperature, int greedy, uint32_t seed_lo, uint32_t seed_hi, int seq_base,
                     int32_t* __restrict__ out_ids, int out_ld, int32_t* __restrict__ cur,
                     const int32_t* __restrict__ forced, int forced_ld, int* __restrict__ pos_ptr,
                     int* __restrict__ step_ptr, float* __restrict__ u_out, float* __restrict__ logits_out) {
    constexpr int E = 256 * VPL;
    __shared__ float sh[E];
    __shared__ float sz[512];
    __shared__ float sred[8];
    const int b = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int step = *step_ptr;
    // ---- ln_f over the row (each thread 2 * VPL elements) ----
    float v[2 * VPL];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * VPL; ++i) {
        v[i] = __bfloat162float(x[static_cast<size_t>(b) * E + i * 128 + tid]);
        sum += v[i];
    }
    sum = warp_sum(sum);
    if (lane == 0) sred[warp] = sum;
    __syncthreads();
    const float mean = (sred[0] + sred[1] + sred[2] + sred[3]) / E;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * VPL; ++i) { const float d = v[i] - mean; sq += d * d; }
    sq = warp_sum(sq);
    if (lane == 0) sred[4 + warp] = sq;
    __syncthreads();
    const float rstd = rsqrtf((sred[4] + sred[5] + sred[6] + sred[7]) / E + eps);
#pragma unroll
    for (int i = 0; i < 2 * VPL; ++i) {
        const int c = i * 128 + tid;
        // the training path rounds ln_f's output to bf16 before the logits GEMM; do the same
        sh[c] = __bfloat162float(__float2bfloat16_rn((v[i] - mean) * rstd * gamma[c] + beta[c]));
    }
    __syncthreads();
    // ---- logits: warp per vocabulary row ----
    float hreg[VPL][8];
#pragma unroll
    for (int k = 0; k < VPL; ++k)
#pragma unroll
        for (int e = 0; e < 8; ++e) hreg[k][e] = sh[(k * 32 + lane) * 8 + e];
    constexpr int RU = 4;     // vocabulary rows in flight per warp (the loop is L2-latency bound otherwise)
    for (int row0 = warp; row0 < V; row0 += 4 * RU) {
        uint4 w[RU][VPL];
#pragma unroll
        for (int r = 0; r < RU; ++r) {
            const int row = min(row0 + 4 * r, V - 1);
#pragma unroll
            for (int k = 0; k < VPL; ++k)
                w[r][k] = __ldg(reinterpret_cast<const uint4*>(wte + static_cast<size_t>(row) * E + (k * 32 + lane) * 8));
        }
#pragma unroll
        for (int r = 0; r < RU; ++r) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < VPL; ++k) acc += dot8(w[r][k], hreg[k]);
            acc = warp_sum(acc);
            if (lane == 0 && row0 + 4 * r < V) sz[row0 + 4 * r] = acc;
        }
    }
    __syncthreads();
    if (logits_out != nullptr)
        for (int c = tid; c < V; c += 128) logits_out[static_cast<size_t>(b) * V + c] = sz[c];
    if (warp == 0) {
        float u;
        int chosen = sample_row(sz, V, inv_temperature, greedy, seed_lo, seed_hi, static_cast<uint32_t>(seq_base + b),
                                static_cast<uint32_t>(step), lane, &u);
        if (forced != nullptr) {
            const int f = forced[static_cast<size_t>(b) * forced_ld + step];
            if (f >= 0) chosen = f;
        }
        if (lane == 0) {
            out_ids[static_cast<size_t>(b) * out_ld + step] = chosen;
            cur[b] = chosen;
            if (u_out != nullptr) u_out[static_cast<size_t>(b) * out_ld + step] = u;
        }
    }
    advance_counters(pos_ptr, step_ptr, step);
}

int logits_sample(const __nv_bfloat16* x, const float* gamma, const float* beta, float eps, const __nv_bfloat16* wte,
                  int E, int V, float temperature, uint64_t seed, int seq_base, int32_t* out_ids, int out_ld,
                  int32_t* cur, const int32_t* forced, int forced_ld, int* pos_ptr, int* step_ptr, float* u_out,
                  float* logits_out, int B, cudaStream_t s) {
    if (B == 0) return 0;
    CB200_REQUIRE(V <= 512 && E % 256 == 0 && E <= 1024, "logits_sample supports vocab <= 512 and E in {256, 512, 768, 1024}");
    const int greedy = temperature <= 0.f ? 1 : 0;
    const float inv_t = greedy ? 1.f : 1.0f / temperature;
    const uint32_t lo = static_cast<uint32_t>(seed), hi = static_cast<uint32_t>(seed >> 32);
    switch (E / 256) {
        case 1: logits_sample_kernel<1><<<B, 128, 0, s>>>(x, gamma, beta, eps, wte, V, inv_t, greedy, lo, hi, seq_base, out_ids, out_ld, cur, forced, forced_ld, pos_ptr, step_ptr, u_out, logits_out); break;
        case 2: logits_sample_kernel<2><<<B, 128, 0, s>>>(x, gamma, beta, eps, wte, V, inv_t, greedy, lo, hi, seq_base, out_ids, out_ld, cur, forced, forced_ld, pos_ptr, step_ptr, u_out, logits_out); break;
        case 3: logits_sample_kernel<3><<<B, 128, 0, s>>>(x, gamma, beta, eps, wte, V, inv_t, greedy, lo, hi, seq_base, out_ids, out_ld, cur, forced, forced_ld, pos_ptr, step_ptr, u_out, logits_out); break;
        default: logits_sample_kernel<4><<<B, 128, 0, s>>>(x, gamma, beta, eps, wte, V, inv_t, greedy, lo, hi, seq_base, out_ids, out_ld, cur, forced, forced_ld, pos_ptr, step_ptr, u_out, logits_out); break;
    }
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

int sample_tokens(const float* logits, int ld, int V, float temperature, uint64_t seed, int seq_base, int32_t* out_ids,
                  int out_ld, int32_t* cur, const int32_t* forced, int forced_ld, int* pos_ptr, int* step_ptr,
                  float* u_out, int B, cudaStream_t s) {
    if (B == 0) return 0;
    const int greedy = temperature <= 0.f ? 1 : 0;
    const float inv_t = greedy ? 1.f : 1.0f / temperature;
    sample_kernel<<<(B + 3) / 4, 128, 0, s>>>(logits, ld, V, inv_t, greedy, static_cast<uint32_t>(seed),
                                              static_cast<uint32_t>(seed >> 32), seq_base, out_ids, out_ld, cur, forced,
                                              forced_ld, pos_ptr, step_ptr, u_out, B);
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

// ---------------------------------------------------------------------------
// Skinny linear layer for decoding: Y[B, N] = X[B, K] W + b with B <= a few
// hundred rows.  The persistent tcgen05 GEMM tiles 128 x 256 outputs per CTA,
// which leaves 2-8 CTAs busy at these sizes; here a CTA owns 128 rows x NB (16
// or 32) output columns, so a layer spreads over N/NB x ceil(B/128) CTAs (48-64
// for c_attn / c_fc), streams K in 64-wide chunks through a 3-stage cp.async
// pipeline and uses mma.sync (weights are L2-resident: the step is latency-,
// not FLOP-bound).  W is the [N, K] ("out, in") bf16 shadow.
// Epilogues: 0 bias, 1 bias + gelu, 2 bias + residual.
// ---------------------------------------------------------------------------
constexpr int DL_THREADS = 256;
constexpr int DL_STAGES = 3;

__device__ __forceinline__ uint32_t dl_tile_off(int row, int chunk) {   // rows of 64 bf16 = 128 bytes, 8 chunks
    return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}

template <int NB, int EPI>
__global__ void __launch_bounds__(DL_THREADS)
decode_linear_kernel(const __nv_bfloat16* __restrict__ X, int ldx, const __nv_bfloat16* __restrict__ Wt,
                     const float* __restrict__ bias, const __nv_bfloat16* __restrict__ res, int ldres,
                     __nv_bfloat16* __restrict__ Y, int ldy, int B, int N, int K) {
    constexpr int X_BYTES = 128 * 128;          // 128 rows x 64 k
    constexpr int W_BYTES = NB * 128;
    constexpr int STAGE = X_BYTES + W_BYTES;
    extern __shared__ __align__(128) uint8_t dl_smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int g = lane >> 2, tig = lane & 3;
    const int n0 = blockIdx.x * NB, r0 = blockIdx.y * 128;
    const int kchunks = K / 64;

    auto issue = [&](int kc, int stage) {
        const uint32_t sx = smem_u32(dl_smem + stage * STAGE), sw = sx + X_BYTES;
#pragma unroll
        for (int c = 0; c < 4; ++c) {           // 128 rows x 8 chunks / 256 threads
            const int idx = tid + c * DL_THREADS;
            const int r = idx >> 3, ch = idx & 7;
            const bool ok = (r0 + r) < B;
            cp_async_16(sx + dl_tile_off(r, ch), X + static_cast<size_t>(ok ? r0 + r : 0) * ldx + kc * 64 + ch * 8, ok);
        }
        if (tid < NB * 8) {
            const int r = tid >> 3, ch = tid & 7;
            const bool ok = (n0 + r) < N;
            cp_async_16(sw + dl_tile_off(r, ch), Wt + static_cast<size_t>(ok ? n0 + r : 0) * K + kc * 64 + ch * 8, ok);
        }
    };

    float acc[NB / 8][4];
#pragma unroll
    for (int t = 0; t < NB / 8; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }

#pragma unroll
    for (int s = 0; s < DL_STAGES - 1; ++s) {
        if (s < kchunks) issue(s, s);
        cp_async_commit();
    }
    const int a_row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, a_chunk = lane >> 4;
    const int b_row = ((lane >> 4) << 3) + (lane & 7), b_chunk = (lane >> 3) & 1;
    for (int kc = 0; kc < kchunks; ++kc) {
        cp_async_wait<DL_STAGES - 2>();
        __syncthreads();
        if (kc + DL_STAGES - 1 < kchunks) issue(kc + DL_STAGES - 1, (kc + DL_STAGES - 1) % DL_STAGES);
        cp_async_commit();
        const uint32_t sx = smem_u32(dl_smem + (kc % DL_STAGES) * STAGE), sw = sx + X_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t af[4];
            ldmatrix_x4(af, sx + dl_tile_off(a_row, ks * 2 + a_chunk));
#pragma unroll
            for (int tp = 0; tp < NB / 16; ++tp) {
                uint32_t bf[4];
                ldmatrix_x4(bf, sw + dl_tile_off(tp * 16 + b_row, ks * 2 + b_chunk));
                mma_bf16_16816(acc[2 * tp], af, bf[0], bf[1]);
                mma_bf16_16816(acc[2 * tp + 1], af, bf[2], bf[3]);
            }
        }
    }
    // ---- epilogue ----
    const int row_lo = r0 + warp * 16 + g, row_hi = row_lo + 8;
#pragma unroll
    for (int t = 0; t < NB / 8; ++t) {
        const int col = n0 + t * 8 + 2 * tig;
        if (col >= N) continue;
        const float b0 = bias ? bias[col] : 0.f, b1 = bias ? bias[col + 1] : 0.f;
        float v[4] = {acc[t][0] + b0, acc[t][1] + b1, acc[t][2] + b0, acc[t][3] + b1};
        if (EPI == 1) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = gelu_tanh(v[e]);
        }
        if (EPI == 2) {
            if (row_lo < B) {
                const float2 r = unpack_bf16(*reinterpret_cast<const uint32_t*>(res + static_cast<size_t>(row_lo) * ldres + col));
                v[0] += r.x; v[1] += r.y;
            }
            if (row_hi < B) {
                const float2 r = unpack_bf16(*reinterpret_cast<const uint32_t*>(res + static_cast<size_t>(row_hi) * ldres + col));
                v[2] += r.x; v[3] += r.y;
            }
        }
        if (row_lo < B) *reinterpret_cast<uint32_t*>(Y + static_cast<size_t>(row_lo) * ldy + col) = pack_bf16(v[0], v[1]);
        if (row_hi < B) *reinterpret_cast<uint32_t*>(Y + static_cast<size_t>(row_hi) * ldy + col) = pack_bf16(v[2], v[3]);
    }
}

template <int NB, int EPI>
static int launch_decode_linear(const __nv_bfloat16* X, int ldx, const __nv_bfloat16* Wt, const float* bias,
                                const __nv_bfloat16* res, int ldres, __nv_bfloat16* Y, int ldy, int B, int N, int K,
                                cudaStream_t s) {
    constexpr size_t smem = DL_STAGES * (128 * 128 + NB * 128);
    auto kernel = decode_linear_kernel<NB, EPI>;
    static bool configured = false;
    if (!configured) {
        CB200_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid((N + NB - 1) / NB, (B + 127) / 128);
    kernel<<<grid, DL_THREADS, smem, s>>>(X, ldx, Wt, bias, res, ldres, Y, ldy, B, N, K);
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

int decode_linear(int epilogue, const __nv_bfloat16* X, int ldx, const __nv_bfloat16* Wt, const float* bias,
                  const __nv_bfloat16* res, int ldres, __nv_bfloat16* Y, int ldy, int B, int N, int K, cudaStream_t s) {
    if (B == 0) return 0;
    CB200_REQUIRE(K % 64 == 0 && N % 8 == 0, "decode_linear needs K %% 64 == 0 and N %% 8 == 0");
    CB200_REQUIRE(epilogue != 2 || res != nullptr, "residual epilogue needs a residual");
    const bool narrow = N <= 512;    // spread small layers over more CTAs
    switch (epilogue) {
        case 0: return narrow ? launch_decode_linear<16, 0>(X, ldx, Wt, bias, res, ldres, Y, ldy, B, N, K, s)
                              : launch_decode_linear<32, 0>(X, ldx, Wt, bias, res, ldres, Y, ldy, B, N, K, s);
        case 1: return narrow ? launch_decode_linear<16, 1>(X, ldx, Wt, bias, res, ldres, Y, ldy, B, N, K, s)
                              : launch_decode_linear<32, 1>(X, ldx, Wt, bias, res, ldres, Y, ldy, B, N, K, s);
        case 2: return narrow ? launch_decode_linear<16, 2>(X, ldx, Wt, bias, res, ldres, Y, ldy, B, N, K, s)
                              : launch_decode_linear<32, 2>(X, ldx, Wt, bias, res, ldres, Y, ldy, B, N, K, s);
        default: break;
    }
    set_error("unknown decode_linear epilogue %d", epilogue);
    return -1;
}

}  // namespace cb200
