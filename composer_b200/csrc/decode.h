// Internal (C++) interface of the decoding kernels (decode.cu).
#pragma once

#include "common.cuh"

namespace cb200 {

// qkv: [B, 3E] bf16 row of the new token; kcache/vcache: [B, H, t_max, D] bf16 (one layer);
// out: [B, E] bf16; *pos_ptr = position of the new token (device memory).
int decode_attention(const __nv_bfloat16* qkv, __nv_bfloat16* kcache, __nv_bfloat16* vcache, __nv_bfloat16* out,
                     const int* pos_ptr, int B, int H, int D, int t_max, float scale, cudaStream_t s);
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

}  // namespace cb200
