// Autoregressive decoding kernels: single-query attention over an in-place
// KV cache and the fused temperature / softmax / Philox multinomial sampler.
//
// Reference: the model's `past=` path (composer/models/transformer.py:423-437,
// :735-770) re-allocates and copies the whole cache every step (tf.concat); here
// the cache is a preallocated [L][2][B][H][T_max][D] bf16 tensor appended in
// place.  Sampling replaces cli.py:670-673 (logits / temperature,
// tf.random.categorical, last position).
#include "decode.h"

namespace cb200 {

__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ float dot8(const uint4& a, const float (&q)[8]) {
    const float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
    return a0.x * q[0] + a0.y * q[1] + a1.x * q[2] + a1.y * q[3] + a2.x * q[4] + a2.y * q[5] + a3.x * q[6] + a3.y * q[7];
}

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
    for (int key0 = warp * KPW; key0 < n_keys; key0 += 4 * KPW) {
        const int key = key0 + lane / CH;
        const bool valid = key < n_keys;
        uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
        if (valid) {
            // the row appended above was written through the normal path: read it coherently
            if (key == pos) {
                kv = *reinterpret_cast<const uint4*>(kb + static_cast<size_t>(key) * D + part * 8);
                vv = *reinterpret_cast<const uint4*>(vb + static_cast<size_t>(key) * D + part * 8);
            } else {
                kv = ld_nc_v4(kb + static_cast<size_t>(key) * D + part * 8);
                vv = ld_nc_v4(vb + static_cast<size_t>(key) * D + part * 8);
            }
        }
        float s = dot8(kv, q);
#pragma unroll
        for (int o = 1; o < CH; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        s = valid ? s * scale_log2 : -INFINITY;
        const float mn = fmaxf(m, s);
        if (mn != -INFINITY) {
            const float corr = fast_exp2(m - mn);
            const float p = fast_exp2(s - mn);
            l = l * corr + p;
            const float2 v0 = unpack_bf16(vv.x), v1 = unpack_bf16(vv.y), v2 = unpack_bf16(vv.z), v3 = unpack_bf16(vv.w);
            acc[0] = acc[0] * corr + p * v0.x; acc[1] = acc[1] * corr + p * v0.y;
            acc[2] = acc[2] * corr + p * v1.x; acc[3] = acc[3] * corr + p * v1.y;
            acc[4] = acc[4] * corr + p * v2.x; acc[5] = acc[5] * corr + p * v2.y;
            acc[6] = acc[6] * corr + p * v3.x; acc[7] = acc[7] * corr + p * v3.y;
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

int decode_embed(const int32_t* cur, const float* wte, const float* wpe, __nv_bfloat16* out, const int* pos_ptr, int B,
                 int E, int vocab, cudaStream_t s) {
    if (B == 0) return 0;
    decode_embed_kernel<<<(B + 7) / 8, 256, 0, s>>>(cur, wte, wpe, out, pos_ptr, B, E, vocab);
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

// Fused temperature scale + softmax + multinomial draw (one warp per sequence).
// u ~ U[0,1) from Philox4x32-10 keyed by (seed, global sequence index, step), so
// the tokens do not depend on how sequences are sharded over GPUs.  The draw is
// the inverse CDF in vocabulary order.  temperature <= 0 selects argmax (first
// maximum).  When `forced` is non-null and forced[b*forced_ld + step] >= 0 that
// id is emitted instead (prompt teacher-forcing).  The last warp to finish
// advances the position counter for the next graph replay.
__global__ void __launch_bounds__(128)
sample_kernel(const float* __restrict__ logits, int ld, int V, float inv_temperature, int greedy, uint32_t seed_lo,
              uint32_t seed_hi, int seq_base, int32_t* __restrict__ out_ids, int out_ld, int32_t* __restrict__ cur,
              const int32_t* __restrict__ forced, int forced_ld, int* __restrict__ pos_ptr, int* __restrict__ step_ptr,
              float* __restrict__ u_out, int B) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + warp;
    const int step = *step_ptr;
    if (b < B) {
        const float* z = logits + static_cast<size_t>(b) * ld;
        float vmax = -INFINITY;
        int amax = 0x7fffffff;
        for (int c = lane; c < V; c += 32) {
            const float v = z[c];
            if (v > vmax) { vmax = v; amax = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float v2 = __shfl_xor_sync(0xffffffffu, vmax, o);
            const int a2 = __shfl_xor_sync(0xffffffffu, amax, o);
            if (v2 > vmax || (v2 == vmax && a2 < amax)) { vmax = v2; amax = a2; }
        }
        int chosen = amax;
        float u = 0.f;
        if (!greedy) {
            const float kLog2e = 1.4426950408889634f;
            const float c = inv_temperature * kLog2e;
            // pass 1: total mass; lane owns the contiguous slice [lo, hi) so that the CDF is in vocabulary order
            const int per = (V + 31) / 32;
            const int lo = lane * per, hi = min(V, lo + per);
            float mass = 0.f;
            for (int i = lo; i < hi; ++i) mass += exp2f((z[i] - vmax) * c);
            float prefix = mass;   // inclusive scan over lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float t = __shfl_up_sync(0xffffffffu, prefix, o);
                if (lane >= o) prefix += t;
            }
            const float total = __shfl_sync(0xffffffffu, prefix, 31);
            const Philox4 r = philox4x32_10(static_cast<uint32_t>(step), static_cast<uint32_t>(seq_base + b), 0x5A17u, 0u,
                                            seed_lo, seed_hi);
            u = (r.x >> 8) * (1.0f / 16777216.0f);          // 24-bit uniform in [0, 1)
            const float target = u * total;
            // the lane whose slice [prefix - mass, prefix) contains the target resolves the id
            const float before = prefix - mass;
            const bool mine = (target >= before && target < prefix) || (lane == 31 && target >= prefix);
            int pick = -1;
            if (mine) {
                float run = before;
                pick = max(hi - 1, lo);
                for (int i = lo; i < hi; ++i) {
                    run += exp2f((z[i] - vmax) * c);
                    if (target < run) { pick = i; break; }
                }
                if (pick >= V) pick = V - 1;
            }
            // lowest lane that claims wins (slices are disjoint; this only breaks float ties)
            const uint32_t ballot = __ballot_sync(0xffffffffu, pick >= 0);
            const int src = ballot ? (__ffs(ballot) - 1) : 0;
            chosen = __shfl_sync(0xffffffffu, pick, src);
            if (chosen < 0) chosen = amax;
        }
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
    // advance the device-side counters once per launch
    __shared__ int block_done;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int ticket = atomicAdd(reinterpret_cast<unsigned int*>(pos_ptr + 1), 1u);
        block_done = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (block_done && threadIdx.x == 0) {
        pos_ptr[1] = 0;          // reset the ticket
        *pos_ptr = *pos_ptr + 1;
        *step_ptr = step + 1;
    }
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

}  // namespace cb200
