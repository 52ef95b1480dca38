// Internal (C++) interface of the HBM-bound kernels (elementwise.cu).
#pragma once

#include "common.cuh"

namespace cb200 {

struct TransposeJob {
    unsigned long long src_offset;   // element offset into the fp32 parameter arena ([rows, cols] row-major)
    unsigned long long dst_offset;   // element offset into the bf16 destination arena ([cols, dst_ld] row-major)
    int rows, cols, dst_ld, pad;
};

int embed_fwd(const int32_t* ids, const float* wte, const float* wpe, __nv_bfloat16* out, int B, int T, int E,
              int pos0, int vocab, const DropoutParams& drop, cudaStream_t s);
int embed_bwd(const int32_t* ids, const __nv_bfloat16* dh, float* dwte, float* dwpe, int B, int T, int E, int pos0,
              int vocab, const DropoutParams& drop, cudaStream_t s);
int layernorm_fwd(const __nv_bfloat16* x, const float* gamma, const float* beta, __nv_bfloat16* y, float* stats,
                  int rows, int E, float eps, cudaStream_t s);
// Optional tail of layernorm_bwd: what bias_grad would do next with dx as its input, i.e. g = dropout_bwd(dx)
// (written to g_out when dropout is on) and dbias += column sums of g, without a second pass over dx.
struct LnBwdTail {
    __nv_bfloat16* g_out;     // [rows, E]; may be null when drop.threshold16 == 0
    float* dbias;             // [E]; null = no tail
    DropoutParams drop;
    uint32_t site, layer;
};
int layernorm_bwd_tail(const __nv_bfloat16* dy_a, const __nv_bfloat16* dy_b, const __nv_bfloat16* x, const float* stats,
                       const float* gamma, const __nv_bfloat16* dres, __nv_bfloat16* dx, float* dgamma, float* dbeta,
                       int rows, int E, const LnBwdTail& tail, cudaStream_t s);
int layernorm_bwd(const __nv_bfloat16* dy_a, const __nv_bfloat16* dy_b, const __nv_bfloat16* x, const float* stats,
                  const float* gamma, const __nv_bfloat16* dres, __nv_bfloat16* dx, float* dgamma, float* dbeta,
                  int rows, int E, cudaStream_t s);
int bias_grad(const __nv_bfloat16* dy, __nv_bfloat16* g_out, float* dbias, int rows, int N, const DropoutParams& drop,
              uint32_t site, uint32_t layer, cudaStream_t s);
int adam_step(float* p, const float* g, float* m, float* v, __nv_bfloat16* shadow, size_t n, float lr_t, float b1,
              float b2, float eps, float grad_scale, cudaStream_t s);
int cast_bf16(const float* src, __nv_bfloat16* dst, size_t n, cudaStream_t s);
int transpose_cast(const float* params, __nv_bfloat16* dst_base, const TransposeJob* jobs_dev, int num_jobs,
                   cudaStream_t s);
int device_sm_count_ew();

}  // namespace cb200
