// Internal (C++) interface of the decoding kernels (decode.cu).
#pragma once

#include "common.cuh"

namespace cb200 {

// qkv: [B, 3E] bf16 row of the new token; kcache/vcache: [B, H, t_max, D] bf16 (one layer);
// out: [B, E] bf16; *pos_ptr = position of the new token (device memory).
int decode_attention(const __nv_bfloat16* qkv, __nv_bfloat16* kcache, __nv_bfloat16* vcache, __nv_bfloat16* out,
                     const int* pos_ptr, int B, int H, int D, int t_max, float scale, cudaStream_t s);
// Batched prefill: k, v rows of T tokens per sequence from a c_attn output [B * T, 3E] into one layer of the cache.
int kv_export(const __nv_bfloat16* qkv, __nv_bfloat16* kcache, __nv_bfloat16* vcache, int B, int T, int H, int D,
              int t_max, cudaStream_t s);
int decode_embed(const int32_t* cur, const float* wte, const float* wpe, __nv_bfloat16* out, const int* pos_ptr, int B,
                 int E, int vocab, cudaStream_t s);
// pos_ptr[0] = position, pos_ptr[1] = ticket scratch (must be 0); step_ptr = output column.
int sample_tokens(const float* logits, int ld, int V, float temperature, uint64_t seed, int seq_base, int32_t* out_ids,
                  int out_ld, int32_t* cur, const int32_t* forced, int forced_ld, int* pos_ptr, int* step_ptr,
                  float* u_out, int B, cudaStream_t s);

// ln_f + tied logits + sampling fused (one CTA per sequence); logits_out (fp32 [B, V]) may be null.
int logits_sample(const __nv_bfloat16* x, const float* gamma, const float* beta, float eps, const __nv_bfloat16* wte,
                  int E, int V, float temperature, uint64_t seed, int seq_base, int32_t* out_ids, int out_ld,
                  int32_t* cur, const int32_t* forced, int forced_ld, int* pos_ptr, int* step_ptr, float* u_out,
                  float* logits_out, int B, cudaStream_t s);
// Y[B, N] = X[B, K] Wt^T + bias ; epilogue 0 bias, 1 bias + gelu, 2 bias + residual.  Wt is [N, K] bf16.
int decode_linear(int epilogue, const __nv_bfloat16* X, int ldx, const __nv_bfloat16* Wt, const float* bias,
                  const __nv_bfloat16* res, int ldres, __nv_bfloat16* Y, int ldy, int B, int N, int K, cudaStream_t s);

// ---- persistent cluster decode (decode_mega.cu): all steps and layers of a generation in one launch ----
constexpr int MG_MAX_LAYERS = 32;
struct MegaLayer {      // element offsets: LayerNorm / bias into the fp32 parameter arena, weights into the bf16 shadow ([out, in])
    uint32_t ln1_g, ln1_b, attn_b, proj_b, ln2_g, ln2_b, fc_b, proj2_b;
    uint32_t attn_w, proj_w, fc_w, proj2_w;
};
struct MegaArgs {
    const float* params;
    const __nv_bfloat16* shadow;
    __nv_bfloat16* cache;          // [L, B, H, t_max, 2, D]: the k row and the v row of a token are adjacent
    const uint8_t* wstream;        // packed weight stream (set by decode_mega)
    const int32_t* first;          // [B] token fed at step 0
    const int32_t* forced;         // [B, steps] teacher-forced ids (>= 0) or -1; may be null
    int32_t* out_ids;              // [B, steps]
    float* uniforms;               // [B, steps] or null
    float* logits_out;             // [B, V] logits of the last step, or null
    long long* prof;               // 64 cycle counters (phase profile of cluster 0: 15 phases, ring waits per phase from 16, marks inside the linear phases from 24), or null
    long long layer_stride;        // elements per layer of the cache
    int B, E, H, F, V, L, t_max, steps, use_ln, greedy, seq_base;
    int step0;                     // first step to run: positions 0 .. step0-1 are already in the cache (batched prefill)
    int kv_split_log2;             // a pair's KV chunks go to up to 1 << this warps when the CTA has few pairs (set by decode_mega)
    int kv_prefetch;               // KV chunks per warp and attention phase requested into L2 during the GEMM phases (set by decode_mega)
    int async_gather;              // 1: all-gathers complete on mbarriers (st.async), 0: at cluster barriers (set by decode_mega)
    int l2_hints;                  // 1: cache reads evict-first, weight stream evict-last (set by decode_mega)
    float eps, scale_log2, inv_temperature;
    uint32_t seed_lo, seed_hi;
    uint32_t wte, wpe, lnf_g, lnf_b;   // offsets into params
    uint32_t wte_sh;                   // bf16 wte [V, E] in the shadow
    MegaLayer layers[MG_MAX_LAYERS];
};
bool decode_mega_supported(int E, int H, int D, int V, int L);
int decode_mega_capacity(int E, int H, int V, int D, int L, int cluster_size);
// max_clusters > 0 caps the number of clusters (tests); 0 = as many as are co-resident.
// cluster_size: 0 = automatic, 4 or 8 CTAs per cluster.
// stream_ws: device scratch of decode_mega_stream_bytes() bytes for the packed weight stream.
int64_t decode_mega_stream_bytes(int E, int H, int D, int V, int L);
// Batched prefill for the cluster kernel's cache layout ([B, H, t_max, 2, D] per layer, swizzled 16-byte pieces).
int kv_export_mega(const __nv_bfloat16* qkv, __nv_bfloat16* cache_layer, int B, int T, int H, int D, int t_max,
                   cudaStream_t s);
int decode_mega(MegaArgs args, int D, int max_clusters, int cluster_size, uint8_t* stream_ws, int64_t stream_ws_bytes,
                cudaStream_t s);

}  // namespace cb200
