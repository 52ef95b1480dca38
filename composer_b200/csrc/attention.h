// Internal (C++) interface of the attention kernels (attention.cu).
#pragma once

#include "common.cuh"

namespace cb200 {

// qkv: [B, T, 3E] bf16 (c_attn output; q | k | v thirds, head h at columns h*D).
// out: [B, T, E] bf16 (heads merged).  lse: [B, H, T] fp32, log2 domain.
int attention_fwd(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse, int B, int T, int H, int D, float scale,
                  const DropoutParams& drop, uint32_t layer, cudaStream_t s);

// delta: [B, H, T] fp32 scratch.  dq_acc: [B, T, E] fp32, must be zero on entry and is zero again on exit.
// dqkv: [B, T, 3E] bf16 gradient of c_attn's output.
int attention_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* out, const __nv_bfloat16* dout, const float* lse,
                  float* delta, float* dq_acc, __nv_bfloat16* dqkv, int B, int T, int H, int D, float scale,
                  const DropoutParams& drop, uint32_t layer, cudaStream_t s);

// Selects the forward kernel: 0 = tcgen05 / TMEM with P kept in TMEM (default), 1 = round-1 warp-level mma.sync
// kernel (A/B reference), 2 = tcgen05 / TMEM with P staged through shared memory, 3 / 4 = tile-shape variants of 0
// for d_h 16 (64-key tiles, 3 / 4 CTAs per SM).
void attention_set_fwd_impl(int impl);

// tcgen05 / TMEM forward (attention_fwd_tc.cu).
int attention_fwd_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse, int B, int T, int H, int D, float scale,
                     const AttnDropKey& key, int variant, cudaStream_t s);

// tcgen05 / TMEM implementation of the main backward kernel (attention_tc.cu).
int attention_bwd_tc_main(const __nv_bfloat16* qkv, const __nv_bfloat16* dout, const float* lse, const float* delta,
                          float* dq_acc, __nv_bfloat16* dqkv, int B, int T, int H, int D, float scale,
                          const AttnDropKey& key, cudaStream_t s);

// Diagnostic: device buffer (12 x 512 int64) that receives the event timeline of the warps of one backward CTA.
void attention_set_trace(long long* buffer);
long long* attention_get_trace();

int attention_mask_export(uint8_t* mask, int B, int T, int H, const DropoutParams& drop, uint32_t layer, cudaStream_t s);

}  // namespace cb200
