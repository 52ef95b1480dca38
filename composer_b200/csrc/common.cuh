// Shared device/host helpers for the composer_b200 kernels (sm_100a only).
//
// Everything here is plain CUDA C++ plus inline PTX: mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (TMEM alloc / mma / ld / commit), Philox
// and small bf16 utilities.  No CUTLASS/CuTe types are used so that the
// descriptors and barriers are visible at the call sites.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cb200 {

// ---------------------------------------------------------------------------
// Host-side error plumbing (thread-local message returned by cb200_last_error)
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
// Counts kernel launches issued by this library (reported by bench.py as gpu_launches).
void note_launch(long long n = 1);

#define CB200_CUDA_OK(expr)                                                        \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) {                                                   \
            cb200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,         \
                             cudaGetErrorString(_e));                              \
            return -2;                                                             \
        }                                                                          \
    } while (0)

#define CB200_REQUIRE(cond, ...)                                                   \
    do {                                                                           \
        if (!(cond)) {                                                             \
            cb200::set_error(__VA_ARGS__);                                         \
            return -1;                                                             \
        }                                                                          \
    } while (0)

// ---------------------------------------------------------------------------
// Small device utilities
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() {
    uint32_t l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float2 unpack_bf16(uint32_t v) {
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&v);
    return __bfloat1622float2(b);
}

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float fast_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// GELU, tanh form (reference: composer/models/transformer.py:35-40), with the hardware tanh
// (MUFU.TANH, relative error ~2^-11: below the bf16 rounding of the stored result).
__device__ __forceinline__ float gelu_tanh(float x) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    const float t = fast_tanh(k0 * x * fmaf(k1, x * x, 1.0f));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}

// d/dx of the tanh-form GELU.
__device__ __forceinline__ float gelu_tanh_grad(float x) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    const float x2 = x * x;
    const float t = fast_tanh(k0 * x * fmaf(k1, x2, 1.0f));
    const float dt = (1.0f - t * t) * k0 * fmaf(3.0f * k1, x2, 1.0f);
    return fmaf(0.5f * x, dt, fmaf(0.5f, t, 0.5f));
}

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  counter = 128 bits, key = 64 bits.
// ---------------------------------------------------------------------------
struct Philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = static_cast<uint64_t>(M0) * c0;
        uint64_t p1 = static_cast<uint64_t>(M1) * c2;
        uint32_t hi0 = static_cast<uint32_t>(p0 >> 32), lo0 = static_cast<uint32_t>(p0);
        uint32_t hi1 = static_cast<uint32_t>(p1 >> 32), lo1 = static_cast<uint32_t>(p1);
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// Dropout sites; mixed into the Philox key so that streams never collide.
enum DropSite : uint32_t { SITE_EMBD = 1, SITE_ATTN_W = 2, SITE_ATTN_RESID = 3, SITE_MLP = 4 };

struct DropoutParams {
    uint32_t seed_lo, seed_hi;   // user seed
    uint32_t step;               // global step (so masks differ every step)
    uint32_t threshold16;        // drop when 16-bit draw < threshold16 ; 0 = dropout off
    float keep_scale;            // 1 / (1 - rate)
    float rate;                  // the configured rate (0 when off)
};

__host__ __device__ __forceinline__ uint32_t drop_key1(const DropoutParams& p, uint32_t site, uint32_t layer) {
    return p.seed_hi ^ (p.step * 0x9E3779B1u) ^ (site << 28) ^ (layer << 20);
}

// Row-major [rows, cols] dropout: one Philox call yields the keep bits of 8
// consecutive elements (col8 = col / 8) of one row; 16 random bits each.
__host__ __device__ __forceinline__ Philox4 drop_bits_rowmajor(const DropoutParams& p, uint32_t site, uint32_t layer,
                                                               uint32_t row, uint32_t col8) {
    return philox4x32_10(row, col8, 0x0D0Du, site, p.seed_lo, drop_key1(p, site, layer));
}

// Attention-probability dropout.  The T x T keep mask of one (batch, head) pair `bh` is defined per
// (query row i, block of 128 keys jb) by one multiplicative congruential stream modulo 2^32:
//     x_0 = attn_row_seed(base(bh), i, jb) (odd),   x_{p+1} = A x_p,   p = 0 .. 63
// and pair p decides keys 128 jb + 2p and 128 jb + 2p + 1 from two odd multiples of the state:
//     key 2p     is kept iff  float_bits(A x_p) >= thr  or the comparison is unordered  (set.geu.f32)
//     key 2p + 1 is kept iff  float_bits(C x_p) >= thr  or unordered.
// Both words are bijections of the state, hence uniform over the odd 32-bit patterns (the HIGH half of the
// 64-bit product A x is not: it never exceeds A), and the comparison is dominated by their high bits, the good
// ones of a power-of-two MCG.  Reading the random words as fp32 makes the keep decision ONE instruction per
// element that already yields the multiplier 1.0f / 0.0f (FSET.BF): 2 instructions per element in all (IMAD +
// FSET), regenerated identically by forward, backward and the mask export.  A thread that owns a contiguous,
// even-aligned range of a row's keys starts at x_p = x_0 A^p (one multiply).  `thr` is a negative float chosen so
// that the number of 32-bit patterns that compare >= thr (or are NaN) is (1 - rate) 2^32
// (attn_drop_threshold_bits).  The base is drawn once per launch from Philox4x32-10 of (seed, step, layer).
struct AttnDropKey {
    uint32_t k0, k1;
    uint32_t thr_bits;      // fp32 bit pattern of the threshold; 0 = dropout off
    float keep_scale;
};

constexpr uint32_t ATTN_MCG_A = 0x93D765DDu;   // Steele & Vigna (2021), good 32-bit MCG multipliers
constexpr uint32_t ATTN_MCG_C = 0xADB4A92Du;

__host__ __device__ constexpr uint32_t mcg_mul_pow(int n) {
    uint32_t a = 1;
    for (int i = 0; i < n; ++i) a *= ATTN_MCG_A;
    return a;
}

__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

__host__ __device__ __forceinline__ uint32_t attn_drop_base(const AttnDropKey& key, uint32_t bh) {
    return fmix32(bh ^ key.k0) + key.k1;
}

// (row < 2^17, jb < 2^10: injective for T <= 2^17); fmix32 is a bijection, so distinct (row, block) pairs get
// distinct hashes; | 1 makes the state odd (a valid MCG state).
__host__ __device__ __forceinline__ uint32_t attn_row_seed(uint32_t base, uint32_t row, uint32_t jb) {
    return fmix32(base ^ ((row << 10) | jb)) | 1u;
}

// Threshold pattern for a drop rate: with K = round((1 - rate) 2^32) patterns to keep.  Negative threshold
// -f(T): every non-negative pattern (2^31, NaNs included: unordered keeps), the 2^23 - 1 negative NaNs and the
// negative patterns of magnitude <= T are kept, K = 2^31 + 2^23 + T.  Rates above ~0.498 need a positive one.
__host__ __device__ __forceinline__ uint32_t attn_drop_threshold_bits(double rate) {
    if (rate <= 0.0) return 0u;
    const double keep = (1.0 - rate) * 4294967296.0;
    const double edge = 2147483648.0 + 8388608.0;
    if (keep >= edge) {
        double t = keep - edge;
        if (t < 8388608.0) t = 8388608.0;              // stay a normal number
        if (t > 2139095039.0) t = 2139095039.0;        // below infinity
        return 0x80000000u | static_cast<uint32_t>(t);
    }
    double t = edge - 1.0 - keep;                        // keep iff pattern >= +f(T):  K = 2^31 - T + 2^23 - 1
    if (t < 8388608.0) t = 8388608.0;
    if (t > 2139095040.0) t = 2139095040.0;
    return static_cast<uint32_t>(t);
}

__host__ __device__ __forceinline__ AttnDropKey make_attn_drop_key(const DropoutParams& p, uint32_t layer) {
    AttnDropKey k;
    const Philox4 r = philox4x32_10(layer, p.step, SITE_ATTN_W, 0x17u, p.seed_lo, p.seed_hi);
    k.k0 = r.x; k.k1 = r.y;
    if (p.threshold16 == 0) {
        k.thr_bits = 0; k.keep_scale = 1.f;
    } else {
        k.thr_bits = attn_drop_threshold_bits(static_cast<double>(p.rate));
        k.keep_scale = p.keep_scale;
    }
    return k;
}

#ifdef __CUDACC__
// One pair of keep multipliers (1.0f / 0.0f) and the stream advance.
__device__ __forceinline__ void attn_drop_pair(uint32_t& x, float thr, float& m0, float& m1) {
    const uint32_t w0 = x * ATTN_MCG_A, w1 = x * ATTN_MCG_C;
    asm("set.geu.f32.f32 %0, %1, %2;" : "=f"(m0) : "f"(__uint_as_float(w0)), "f"(thr));
    asm("set.geu.f32.f32 %0, %1, %2;" : "=f"(m1) : "f"(__uint_as_float(w1)), "f"(thr));
    x = w0;
}

// ---- packed fp32 pairs (Blackwell: FFMA2 / FADD2 / FMUL2 do two lanes per issue slot) -------------------
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ uint64_t f2_pack(uint32_t lo, uint32_t hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
#endif

// Packed (two lanes per instruction) forms of the GELU pair above for the GEMM epilogues, which are bound by their
// arithmetic (c_fc forward: tensor pipe 28 %): 6 -> 3 FMA-pipe instructions per element, the tanh stays one MUFU each.
__device__ __forceinline__ uint64_t gelu_tanh2(uint64_t x) {
    const uint64_t k0 = f2_pack(0.7978845608028654f, 0.7978845608028654f), k1 = f2_pack(0.044715f, 0.044715f);
    const uint64_t one = f2_pack(1.0f, 1.0f), half = f2_pack(0.5f, 0.5f);
    const uint64_t u = f2_mul(f2_mul(k0, x), f2_fma(k1, f2_mul(x, x), one));
    float u0, u1;
    f2_unpack(u, u0, u1);
    const uint64_t t = f2_pack(fast_tanh(u0), fast_tanh(u1));
    const uint64_t hx = f2_mul(half, x);
    return f2_fma(hx, t, hx);
}
__device__ __forceinline__ uint64_t gelu_tanh_grad2(uint64_t x) {
    const uint64_t k0 = f2_pack(0.7978845608028654f, 0.7978845608028654f), k1 = f2_pack(0.044715f, 0.044715f);
    const uint64_t k3 = f2_pack(3.0f * 0.044715f, 3.0f * 0.044715f);
    const uint64_t one = f2_pack(1.0f, 1.0f), half = f2_pack(0.5f, 0.5f), mone = f2_pack(-1.0f, -1.0f);
    const uint64_t x2 = f2_mul(x, x);
    const uint64_t u = f2_mul(f2_mul(k0, x), f2_fma(k1, x2, one));
    float u0, u1;
    f2_unpack(u, u0, u1);
    const uint64_t t = f2_pack(fast_tanh(u0), fast_tanh(u1));
    // dt = (1 - t^2) k0 (3 k1 x^2 + 1)
    const uint64_t omt2 = f2_fma(f2_mul(mone, t), t, one);
    const uint64_t dt = f2_mul(f2_mul(omt2, k0), f2_fma(k3, x2, one));
    return f2_fma(f2_mul(half, x), dt, f2_fma(half, t, half));
}

__host__ __device__ __forceinline__ uint32_t drop_u16(const Philox4& r, int e) {
    uint32_t w = (e < 2) ? r.x : (e < 4) ? r.y : (e < 6) ? r.z : r.w;
    return (e & 1) ? (w >> 16) : (w & 0xFFFFu);
}

// ---------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Spin on a phase parity.  A bounded spin turns a protocol bug into a trap
// (reported by the runtime as an error) instead of a hung GPU box.  (Loop: see mbar_wait_a below.)
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity);
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { mbar_wait_a(smem_u32(bar), parity); }

// The same operations on 32-bit shared-space addresses.  A kernel that converts its barrier / buffer pointers once
// (smem_u32) and keeps the addresses in registers avoids a generic -> shared conversion (4-6 uniform-datapath
// instructions) at every barrier operation of its inner loops.
__device__ __forceinline__ void mbar_init_a(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// The wait loop is part of every hand-off of the warp-specialised kernels, and their waiting warps share the
// schedulers with the working ones: in the attention backward kernel a third of all executed instructions were wait
// loops (ncu source view: SYNCS + BRA + uniform-datapath bookkeeping).  So the loop is kept minimal, in PTX: a poll
// that succeeds costs two instructions, a failed one five (the hardware parks the thread inside try_wait for a
// system-dependent time), and the spin counter that turns a protocol bug into a trap instead of a hung GPU is only
// touched on the failure path.
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .u32 spins;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MBAR_DONE;\n\t"
        "mov.u32 spins, 0;\n"
        "MBAR_RETRY:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MBAR_DONE;\n\t"
        "add.u32 spins, spins, 1;\n\t"
        "setp.gt.u32 p, spins, 0x1000000;\n\t"
        "@p trap;\n\t"
        "bra MBAR_RETRY;\n"
        "MBAR_DONE:\n\t"
        "}"
        ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_a(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void umma_commit_a(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---------------------------------------------------------------------------
// TMA (2-D tiled tensor maps)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(smem_src)), "r"(x), "r"(y)
                 : "memory");
}

__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }

template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// Make generic-proxy smem writes visible to the async proxy (TMA store).
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 inputs, fp32 accumulate.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Arrive on an mbarrier once all previously issued tcgen05.mma have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets TMEM lane
// (warp_quadrant*32 + t), columns [col, col+32).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(taddr)
                 : "memory");
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Registers -> TMEM: thread t of the warp writes TMEM lane (warp_quadrant*32 + t), columns [col, col+16).
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (M x 16 bf16, K-major) is read from TMEM, where row m
// is lane m and two consecutive k share a 32-bit column (what a 32x32b tcgen05.st of packed pairs writes).
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Register re-distribution between the warpgroups of a CTA (all four warps of the group execute it).
template <int REGS>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }

// ---- UMMA descriptors (cf. PTX ISA "tcgen05 shared memory descriptor") -----
// 64-bit shared-memory matrix descriptor for a SWIZZLE_128B operand tile.
//   bits  0-13 start address >> 4      bits 16-29 leading byte offset >> 4
//   bits 32-45 stride byte offset >> 4 bits 46-47 descriptor version (1 on sm_100)
//   bits 61-63 layout type (2 = SWIZZLE_128B)
//   layout type codes: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(layout_type) << 61;
    return d;
}

__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return umma_smem_desc(smem_addr, lbo_bytes, sbo_bytes, 2u);
}

// The same descriptor as two 32-bit words.  Only the start-address field (bits 0-13 of the low word) changes
// between the MMAs of a loop, so an issuing thread keeps `lo` words of its buffers in registers and adds
// (byte offset >> 4) per k-step: one integer add per operand instead of re-encoding the descriptor (the issuing
// thread executes alone at one instruction every few cycles, and a kernel with many small MMAs per tile is bound by
// that instruction stream, see DESIGN.md "Attention").  No carry leaves the field: shared addresses are < 2^18.
__host__ __device__ constexpr uint32_t umma_desc_hi(uint32_t sbo_bytes, uint32_t layout_type) {
    return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout_type << 29);
}
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
    return ((smem_addr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ void umma_bf16_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand in TMEM (TS form), B descriptor as two words.
__device__ __forceinline__ void umma_bf16_ts_w(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Swizzle mode whose span equals a row of `row_bytes` (32, 64 or 128) bytes.
__host__ __device__ constexpr uint32_t umma_layout_for_row_bytes(int row_bytes) {
    return row_bytes == 128 ? 2u : row_bytes == 64 ? 4u : 6u;
}

// 32-bit instruction descriptor, kind::f16, bf16 x bf16 -> fp32.
//   bits 4-5 D format (1 = f32)   bits 7-9 A format (1 = bf16)   bits 10-12 B format (1 = bf16)
//   bit 15 A major (0 = K, 1 = MN)   bit 16 B major   bits 17-22 N >> 3   bits 24-28 M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
           (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(m >> 4) << 24);
}

}  // namespace cb200
