// Internal (C++) interface of the GEMM family; the C ABI lives in
// include/composer_b200.h and capi.cu.
#pragma once

#include "common.cuh"

namespace cb200 {

enum GemmKind : int {
    GEMM_BIAS = 0,           // out0 = bf16(A B^T + bias)                      A, B K-major
    GEMM_BIAS_GELU = 1,      // out0 = bf16(pre), out1 = bf16(gelu(pre))       A, B K-major
    GEMM_BIAS_DROP_RES = 2,  // out0 = bf16(aux + dropout(A B^T + bias))       A, B K-major
    GEMM_MUL_DGELU = 3,      // out0 = bf16((A B^T) * gelu'(aux))              A, B K-major
    GEMM_WGRAD = 4,          // outf += A^T B   (A, B stored [K, M], [K, N])    split-K fp32 atomics
    GEMM_CE = 5,             // fused softmax cross-entropy over N <= 512       A, B K-major
    GEMM_BIAS_BMN = 6,       // like GEMM_BIAS with B stored [K, N]
};

// D[M, N] = sum_k A(m, k) B(n, k).  K-major operands are stored [M, K] / [N, K]
// (row-major, ld = elements between rows); MN-major operands are stored
// [K, M] / [K, N].
struct GemmDesc {
    GemmKind kind;
    int M, N, K;
    const __nv_bfloat16* A; int lda;
    const __nv_bfloat16* B; int ldb;
    const float* bias;                 // [N] or nullptr
    __nv_bfloat16* out0; int ld_out0;  // [M, N]
    __nv_bfloat16* out1; int ld_out1;  // [M, N] (GELU only)
    const __nv_bfloat16* aux; int ld_aux;
    float* outf; int ld_outf;
    DropoutParams drop; uint32_t drop_site, drop_layer;
    const int32_t* labels; __nv_bfloat16* dlogits; int ld_dlogits; float grad_scale;
    float* loss_sum; int* correct;
};

int gemm_launch(const GemmDesc& d, cudaStream_t stream);
int device_sm_count();
int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t stride_elems,
                   uint32_t box_inner, uint32_t box_outer);

int make_tmap_bf16_sw(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t stride_elems,
                      uint32_t box_inner, uint32_t box_outer, int swizzle_bytes);

}  // namespace cb200
