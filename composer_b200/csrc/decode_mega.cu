// Persistent cluster decode: the whole generation loop (all steps, all layers)
// in ONE kernel launch.
//
// Sequences never interact while decoding, so the batch is split over thread
// block clusters of 8 CTAs, each cluster owning up to 16 sequences (= the M of
// one m16n8k16 MMA) for the whole generation, and there is no grid-wide
// synchronisation at all: the only barriers are hardware cluster barriers
// (~0.2 us).  Inside a cluster every linear layer is split over the 8 CTAs by
// output columns; c_attn is split by heads, so a CTA computes q, k, v of its own
// H/8 heads, appends k, v to the cache and attends for those heads without any
// exchange.  The activations that the next layer needs in full (attention
// output, x2, gelu output, block output: 16 rows each) are all-gathered by
// writing the CTA's column slice into the shared memory of all 8 peers (DSMEM),
// 4 cluster barriers per decoder block.  Weights (bf16 [out, in] shadows) stream
// from L2 straight into mma.sync B fragments (a lane reads 16 contiguous bytes
// of one output row; the k-order inside a 32-wide step is permuted identically
// for A and B, which a dot product does not notice).  The KV cache is streamed
// by TMA bulk copies (cp.async.bulk + mbarrier) into a private 4-stage ring per
// warp, so ~100 KB of cache reads are in flight per SM without holding
// registers; a warp owns one (sequence, head) pair at a time and keeps the
// softmax online.  The final CTA 0 of each cluster draws the tokens
// (same Philox / inverse-CDF rule as logits_sample_kernel) and broadcasts them.
//
// Reference semantics: Transformer.call with `past=` (composer/models/
// transformer.py:423-437, 583-597, 735-833) and the sampling rule of
// cli.py:663-676.  Rounding points (bf16 LayerNorm outputs, q/k/v, attention
// output, residual stream, gelu output) are those of the per-step kernels in
// decode.cu, which stay as the path for shapes this kernel does not cover.
#include "decode.h"
#include "decode_common.cuh"
#include "mma_sync.cuh"

namespace cb200 {

constexpr int MG_CL = 8;            // CTAs per cluster
constexpr int MG_WARPS = 16;
constexpr int MG_THREADS = MG_WARPS * 32;
constexpr int MG_ROWS = 16;         // sequences per cluster
constexpr int MG_CHUNK = 1024;      // bytes of K (and of V) per ring stage

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, const uint4& v) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ uint4 ldg_nc_v4_keep(const void* p) {   // weights: let L1 keep the other half of the line
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct MegaSmem {        // byte offsets into dynamic shared memory (computed on the host, identical in every CTA)
    int pe, pf;          // row pitch of the [16, E] and [16, F] bf16 buffers (2E + 64, 2F + 64: conflict-free A reads)
    int buf0, buf1, bufn, bufg, qkv, red, ring, bars, toks, total;
    int nst;             // ring stages per warp
    int zp;              // floats per row of the logits matrix (aliases the ring of CTA 0)
};

// A[16, K] (bf16, shared memory, row pitch `pitch` bytes)  x  W[rows of an [N, K] bf16 matrix]^T.
// Work unit = (n-tile of 8 output columns, K split); a warp takes units warp, warp + 16, ...  Partial sums go to
// red[ks][row][ncols] (fp32).  wrow(nt) = first weight row of n-tile nt (rows wrow .. wrow + 7).
template <typename RowFn>
__device__ __forceinline__ void mma_units(const uint8_t* A, int pitch, int K, const __nv_bfloat16* __restrict__ W,
                                          int ntiles, int ksplit, RowFn wrow, int row_limit, float* red, int warp,
                                          int lane) {
    const int g = lane >> 2, tig = lane & 3;
    const int ncols = ntiles * 8;
    const int kper = K / ksplit;
    for (int u = warp; u < ntiles * ksplit; u += MG_WARPS) {
        const int nt = u % ntiles, ks = u / ntiles;
        const int wr_row = min(wrow(nt) + g, row_limit);
        const __nv_bfloat16* wr = W + static_cast<size_t>(wr_row) * K + ks * kper + 8 * tig;
        const uint8_t* a0 = A + g * pitch + (ks * kper + 8 * tig) * 2;
        const uint8_t* a1 = a0 + 8 * pitch;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int kk = 0; kk < kper; kk += 256) {
            uint4 w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (kk + i * 32 < kper) w[i] = ldg_nc_v4_keep(wr + kk + i * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (kk + i * 32 < kper) {
                    const uint4 xa = *reinterpret_cast<const uint4*>(a0 + (kk + i * 32) * 2);
                    const uint4 xb = *reinterpret_cast<const uint4*>(a1 + (kk + i * 32) * 2);
                    mma_16816(acc, xa.x, xb.x, xa.y, xb.y, w[i].x, w[i].y);
                    mma_16816(acc, xa.z, xb.z, xa.w, xb.w, w[i].z, w[i].w);
                }
        }
        float* r = red + (ks * MG_ROWS + g) * ncols + nt * 8 + 2 * tig;
        *reinterpret_cast<float2*>(r) = make_float2(acc[0], acc[1]);
        *reinterpret_cast<float2*>(r + 8 * ncols) = make_float2(acc[2], acc[3]);
    }
}

// Sums the K-split partials of output chunk (row, 8 columns starting at col) and adds the bias.
__device__ __forceinline__ void gather8(const float* red, int ksplit, int ncols, int row, int col, const float* bias,
                                        float (&v)[8]) {
    if (bias != nullptr) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias)), b1 = __ldg(reinterpret_cast<const float4*>(bias + 4));
        v[0] = b0.x; v[1] = b0.y; v[2] = b0.z; v[3] = b0.w; v[4] = b1.x; v[5] = b1.y; v[6] = b1.z; v[7] = b1.w;
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
    }
    for (int ks = 0; ks < ksplit; ++ks) {
        const float* r = red + (ks * MG_ROWS + row) * ncols + col;
        const float4 p0 = *reinterpret_cast<const float4*>(r), p1 = *reinterpret_cast<const float4*>(r + 4);
        v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w; v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
    }
}

__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

__device__ __forceinline__ void add8(float (&v)[8], const uint4& r) {
    const float2 a = unpack_bf16(r.x), b = unpack_bf16(r.y), c = unpack_bf16(r.z), d = unpack_bf16(r.w);
    v[0] += a.x; v[1] += a.y; v[2] += b.x; v[3] += b.y; v[4] += c.x; v[5] += c.y; v[6] += d.x; v[7] += d.y;
}

// Writes a 16-byte chunk to the same shared-memory offset of all 8 CTAs of the cluster.
__device__ __forceinline__ void broadcast16(const uint8_t* local, const uint4& val) {
    const uint32_t a = smem_u32(local);
#pragma unroll
    for (int r = 0; r < MG_CL; ++r) st_cluster_v4(map_to_cta(a, r), val);
}

// LayerNorm of the 16 rows of `src` into `dst` (warp = row), both [16, E] bf16 with pitch pe.  Two passes in
// registers like layernorm_fwd_kernel; the output is rounded to bf16 (what the next GEMM consumes).
__device__ __forceinline__ void layernorm_rows(const uint8_t* src, uint8_t* dst, int pe, int E, const float* gamma,
                                               const float* beta, float eps, bool enabled, int warp, int lane) {
    const uint8_t* s = src + warp * pe;
    uint8_t* d = dst + warp * pe;
    const int nchunk = E / 8;                 // 16-byte chunks per row: 32 (E 256) or 64 (E 512)
    float v[2][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int ch = lane + 32 * i;
        if (ch < nchunk) {
            const uint4 raw = *reinterpret_cast<const uint4*>(s + ch * 16);
            const float2 a = unpack_bf16(raw.x), b = unpack_bf16(raw.y), c = unpack_bf16(raw.z), e = unpack_bf16(raw.w);
            v[i][0] = a.x; v[i][1] = a.y; v[i][2] = b.x; v[i][3] = b.y; v[i][4] = c.x; v[i][5] = c.y; v[i][6] = e.x; v[i][7] = e.y;
#pragma unroll
            for (int k = 0; k < 8; ++k) sum += v[i][k];
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[i][k] = 0.f;
        }
    }
    if (!enabled) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int ch = lane + 32 * i;
            if (ch < nchunk) *reinterpret_cast<uint4*>(d + ch * 16) = *reinterpret_cast<const uint4*>(s + ch * 16);
        }
        return;
    }
    const float mean = warp_sum(sum) / E;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i)
        if (lane + 32 * i < nchunk)
#pragma unroll
            for (int k = 0; k < 8; ++k) { const float t = v[i][k] - mean; sq += t * t; }
    const float rstd = rsqrtf(warp_sum(sq) / E + eps);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int ch = lane + 32 * i;
        if (ch < nchunk) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + ch * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + ch * 8 + 4));
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + ch * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + ch * 8 + 4));
            float o[8];
            o[0] = (v[i][0] - mean) * rstd * g0.x + b0.x; o[1] = (v[i][1] - mean) * rstd * g0.y + b0.y;
            o[2] = (v[i][2] - mean) * rstd * g0.z + b0.z; o[3] = (v[i][3] - mean) * rstd * g0.w + b0.w;
            o[4] = (v[i][4] - mean) * rstd * g1.x + b1.x; o[5] = (v[i][5] - mean) * rstd * g1.y + b1.y;
            o[6] = (v[i][6] - mean) * rstd * g1.z + b1.z; o[7] = (v[i][7] - mean) * rstd * g1.w + b1.w;
            *reinterpret_cast<uint4*>(d + ch * 16) = pack8(o);
        }
    }
}

template <int D>
__global__ void __launch_bounds__(MG_THREADS, 1)
decode_mega_kernel(const __grid_constant__ MegaArgs a, const __grid_constant__ MegaSmem sm) {
    constexpr int CH = D / 8;                 // 16-byte chunks per head row
    constexpr int KPW = 32 / CH;              // keys per warp-wide shared-memory read
    constexpr int CT = MG_CHUNK / (2 * D);    // tokens per ring stage (= 2 * KPW)
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int crank = static_cast<int>(cluster_ctarank());
    const int cid = blockIdx.x / MG_CL, ncl = gridDim.x / MG_CL;
    // sequences of this cluster: the first (B % ncl) clusters take one more
    const int base = a.B / ncl, rem = a.B % ncl;
    const int G = base + (cid < rem ? 1 : 0);
    const int s0 = cid * base + min(cid, rem);
    if (G == 0) return;                        // whole cluster leaves together

    const int E = a.E, F = a.F, H = a.H, V = a.V;
    const int HS = E / MG_CL;                  // columns of the residual stream owned by this CTA (its heads)
    const int FS = F / MG_CL;
    const int HPC = H / MG_CL;                 // heads per CTA
    const int VS = 8 * ((V + 63) / 64);        // vocabulary rows per CTA (whole n-tiles)
    const int pe = sm.pe, pf = sm.pf;
    uint8_t* bufU = smem + sm.buf0;            // block input x, later x2
    uint8_t* bufW = smem + sm.buf1;            // attention output, later the block output
    uint8_t* bufN = smem + sm.bufn;            // LayerNorm output (ln_1: the residual stream of the block)
    uint8_t* bufG = smem + sm.bufg;            // gelu output [16, F]
    uint8_t* qkvs = smem + sm.qkv;             // q | k | v of this CTA's heads, [16][3 * HS] bf16
    float* red = reinterpret_cast<float*>(smem + sm.red);
    uint8_t* ring = smem + sm.ring;
    float* Z = reinterpret_cast<float*>(smem + sm.ring);   // logits [16][zp], valid in CTA 0 between the last two barriers of a step
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sm.bars);
    int* toks = reinterpret_cast<int*>(smem + sm.toks);
    const int NST = sm.nst;

    // ---- one-time setup ----
    for (int i = tid * 16; i < sm.bars; i += MG_THREADS * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    if (tid < MG_WARPS * NST) mbar_init(&bars[tid], 1);
    if (tid < MG_ROWS) toks[tid] = (tid < G) ? a.first[s0 + tid] : 0;
    mbar_fence_init();
    __syncthreads();
    cluster_sync_all();                        // every CTA of the cluster is resident before any DSMEM access

    // optional phase profile (cluster 0, CTA 0, thread 0): cycles accumulated per phase over the whole generation
    long long prof_acc[16];
    long long prof_t = 0;
    const bool profiling = a.prof != nullptr && blockIdx.x == 0 && tid == 0;
    if (profiling) {
#pragma unroll
        for (int i = 0; i < 16; ++i) prof_acc[i] = 0;
        prof_t = clock64();
    }
#define MG_PROF(slot)                                                        \
    if (profiling) {                                                         \
        const long long now_ = clock64();                                    \
        prof_acc[slot] += now_ - prof_t;                                     \
        prof_t = now_;                                                       \
    }
    uint32_t ring_count = 0;                   // stages consumed by this warp so far (stage = count % NST, parity from count / NST)
    const float* P = a.params;
    const __nv_bfloat16* S = a.shadow;

    for (int step = 0; step < a.steps; ++step) {
        const int pos = step;
        // ---- token + positional embedding: warp = row ----
        if (warp < G) {
            int id = toks[warp];
            id = min(max(id, 0), V - 1);
            const float* te = P + a.wte + static_cast<size_t>(id) * E;
            const float* pp = P + a.wpe + static_cast<size_t>(pos) * E;
            for (int ch = lane; ch < E / 8; ch += 32) {
                const float4 t0 = __ldg(reinterpret_cast<const float4*>(te + ch * 8)), t1 = __ldg(reinterpret_cast<const float4*>(te + ch * 8 + 4));
                const float4 p0 = __ldg(reinterpret_cast<const float4*>(pp + ch * 8)), p1 = __ldg(reinterpret_cast<const float4*>(pp + ch * 8 + 4));
                const float o[8] = {t0.x + p0.x, t0.y + p0.y, t0.z + p0.z, t0.w + p0.w, t1.x + p1.x, t1.y + p1.y, t1.z + p1.z, t1.w + p1.w};
                *reinterpret_cast<uint4*>(bufU + warp * pe + ch * 16) = pack8(o);
            }
        }
        __syncthreads();
        MG_PROF(0)

        uint8_t* X = bufU;                     // block input (full rows)
        uint8_t* Y = bufW;                     // the other full-row buffer
        for (int l = 0; l < a.L; ++l) {
            const MegaLayer& lw = a.layers[l];
            // ---- P1: x1 = ln_1(x) ----
            layernorm_rows(X, bufN, pe, E, P + lw.ln1_g, P + lw.ln1_b, a.eps, a.use_ln != 0, warp, lane);
            __syncthreads();
            MG_PROF(1)
            // ---- P2: q, k, v of this CTA's heads ----
            {
                const int ntiles = 3 * HS / 8;
                const int tiles_per_part = HS / 8;
                mma_units(bufN, pe, E, S + lw.attn_w, ntiles, 1,
                          [&](int nt) { return (nt / tiles_per_part) * E + crank * HS + (nt % tiles_per_part) * 8; },
                          3 * E - 1, red, warp, lane);
                __syncthreads();
                const int ncols = 3 * HS, cpr = ncols / 8;
                for (int it = tid; it < MG_ROWS * cpr; it += MG_THREADS) {
                    const int row = it / cpr, col = (it % cpr) * 8;
                    const int part = col / HS, jj = col % HS;
                    float v[8];
                    gather8(red, 1, ncols, row, col, P + lw.attn_b + part * E + crank * HS + jj, v);
                    *reinterpret_cast<uint4*>(qkvs + (row * ncols + col) * 2) = pack8(v);
                }
                __syncthreads();
            }
            MG_PROF(2)
            // ---- P3: append k, v; attention of (sequence, head) pairs; all-gather the output into Y ----
            {
                __nv_bfloat16* kc = a.cache + static_cast<size_t>(l) * a.layer_stride;
                __nv_bfloat16* vc = kc + a.layer_stride / 2;
                const int npairs = G * HPC;
                const int nmine = (npairs > warp) ? (npairs - warp + MG_WARPS - 1) / MG_WARPS : 0;
                const int nchunks = (pos + CT - 1) / CT;
                const int njobs = nmine * nchunks;
                const int part = lane % CH;
                auto issue = [&](int j) {      // lane 0: stream chunk j of this warp's job list into its ring slot
                    const int q = warp + MG_WARPS * (j / nchunks), c = j % nchunks;
                    const int b = s0 + q / HPC, h = crank * HPC + q % HPC;
                    const size_t off = ((static_cast<size_t>(b) * H + h) * a.t_max + static_cast<size_t>(c) * CT) * D;
                    const uint32_t bytes = static_cast<uint32_t>(min(CT, pos - c * CT)) * 2 * D;
                    const uint32_t slot = (ring_count + j) % NST;
                    uint64_t* bar = &bars[warp * NST + slot];
                    uint8_t* dst = ring + (warp * NST + slot) * (2 * MG_CHUNK);
                    mbar_expect_tx(bar, 2 * bytes);
                    bulk_g2s(smem_u32(dst), kc + off, bytes, bar);
                    bulk_g2s(smem_u32(dst + MG_CHUNK), vc + off, bytes, bar);
                };
                if (lane == 0)
                    for (int j = 0; j < min(NST, njobs); ++j) issue(j);
                for (int pi = 0; pi < nmine; ++pi) {
                    const int q = warp + MG_WARPS * pi;
                    const int sl = q / HPC, hh = q % HPC;
                    const int b = s0 + sl, h = crank * HPC + hh;
                    const uint8_t* qrow = qkvs + (sl * 3 * HS + hh * D + part * 8) * 2;
                    float qf[8];
                    {
                        const uint4 qv = *reinterpret_cast<const uint4*>(qrow);
                        const float2 a0 = unpack_bf16(qv.x), a1 = unpack_bf16(qv.y), a2 = unpack_bf16(qv.z), a3 = unpack_bf16(qv.w);
                        qf[0] = a0.x; qf[1] = a0.y; qf[2] = a1.x; qf[3] = a1.y; qf[4] = a2.x; qf[5] = a2.y; qf[6] = a3.x; qf[7] = a3.y;
                    }
                    const uint4 knew = *reinterpret_cast<const uint4*>(qrow + HS * 2);
                    const uint4 vnew = *reinterpret_cast<const uint4*>(qrow + 2 * HS * 2);
                    // append (global cache, read back by TMA in later steps)
                    if (lane < 2 * CH) {
                        const size_t off = ((static_cast<size_t>(b) * H + h) * a.t_max + pos) * D + part * 8;
                        *reinterpret_cast<uint4*>((lane < CH ? kc : vc) + off) = (lane < CH) ? knew : vnew;
                    }
                    // the new token seeds the online softmax of the first lane group
                    float snew = dot8(knew, qf);
#pragma unroll
                    for (int o = 1; o < CH; o <<= 1) snew += __shfl_xor_sync(0xffffffffu, snew, o);
                    float m = -INFINITY, lsum = 0.f, acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    if (lane < CH) {
                        m = snew * a.scale_log2;
                        lsum = 1.f;
                        const float2 v0 = unpack_bf16(vnew.x), v1 = unpack_bf16(vnew.y), v2 = unpack_bf16(vnew.z), v3 = unpack_bf16(vnew.w);
                        acc[0] = v0.x; acc[1] = v0.y; acc[2] = v1.x; acc[3] = v1.y; acc[4] = v2.x; acc[5] = v2.y; acc[6] = v3.x; acc[7] = v3.y;
                    }
                    for (int c = 0; c < nchunks; ++c) {
                        const int j = pi * nchunks + c;
                        const uint32_t cnt = ring_count + j;
                        const uint32_t slot = cnt % NST;
                        mbar_wait(&bars[warp * NST + slot], (cnt / NST) & 1);
                        const uint8_t* sk = ring + (warp * NST + slot) * (2 * MG_CHUNK);
                        const uint8_t* sv = sk + MG_CHUNK;
                        const int ntok = min(CT, pos - c * CT);
                        uint4 kv[2], vv[2];
                        float sc[2];
                        float mn = m;
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int key = u * KPW + lane / CH;
                            const bool valid = key < ntok;
                            kv[u] = *reinterpret_cast<const uint4*>(sk + key * 2 * D + part * 16);
                            vv[u] = *reinterpret_cast<const uint4*>(sv + key * 2 * D + part * 16);
                            if (!valid) vv[u] = make_uint4(0, 0, 0, 0);
                            float s = dot8(kv[u], qf);
#pragma unroll
                            for (int o = 1; o < CH; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                            sc[u] = valid ? s * a.scale_log2 : -INFINITY;
                            mn = fmaxf(mn, sc[u]);
                        }
                        if (mn != -INFINITY) {
                            const float corr = fast_exp2(m - mn);
                            lsum *= corr;
#pragma unroll
                            for (int e = 0; e < 8; ++e) acc[e] *= corr;
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                const float p = fast_exp2(sc[u] - mn);
                                lsum += p;
                                const float2 v0 = unpack_bf16(vv[u].x), v1 = unpack_bf16(vv[u].y), v2 = unpack_bf16(vv[u].z), v3 = unpack_bf16(vv[u].w);
                                acc[0] += p * v0.x; acc[1] += p * v0.y; acc[2] += p * v1.x; acc[3] += p * v1.y;
                                acc[4] += p * v2.x; acc[5] += p * v2.y; acc[6] += p * v3.x; acc[7] += p * v3.y;
                            }
                            m = mn;
                        }
                        __syncwarp();
                        if (lane == 0 && j + NST < njobs) issue(j + NST);
                    }
                    // merge the lane groups (every lane ends with the totals of its slice `part`)
#pragma unroll
                    for (int o = CH; o < 32; o <<= 1) {
                        const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
                        const float l2 = __shfl_xor_sync(0xffffffffu, lsum, o);
                        const float mn = fmaxf(m, m2);
                        const float c1 = (m == -INFINITY) ? 0.f : fast_exp2(m - mn);
                        const float c2 = (m2 == -INFINITY) ? 0.f : fast_exp2(m2 - mn);
                        lsum = lsum * c1 + l2 * c2;
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float a2 = __shfl_xor_sync(0xffffffffu, acc[e], o);
                            acc[e] = acc[e] * c1 + a2 * c2;
                        }
                        m = mn;
                    }
                    const float inv = 1.0f / lsum;
                    float o8[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) o8[e] = acc[e] * inv;
                    const uint4 packed = pack8(o8);
                    const uint32_t dst = smem_u32(Y + sl * pe + (h * D + part * 8) * 2);
                    for (int i = lane; i < MG_CL * CH; i += 32) st_cluster_v4(map_to_cta(dst, i / CH), packed);
                }
                ring_count += njobs;
                // the appended rows are read through the async proxy (TMA) in later steps
                asm volatile("fence.proxy.async.global;" ::: "memory");
            }
            MG_PROF(3)
            cluster_sync_all();                // A: attention output of all heads is in Y everywhere
            MG_PROF(4)
            // ---- P4: x2 = x1 + c_proj(att) for this CTA's columns, all-gathered into X ----
            {
                const int ntiles = HS / 8;
                const int ksplit = max(1, min(MG_WARPS / ntiles, E / 64));
                mma_units(Y, pe, E, S + lw.proj_w, ntiles, ksplit, [&](int nt) { return crank * HS + nt * 8; }, E - 1, red,
                          warp, lane);
                __syncthreads();
                const int cpr = HS / 8;
                for (int it = tid; it < MG_ROWS * cpr; it += MG_THREADS) {
                    const int row = it / cpr, col = (it % cpr) * 8, gcol = crank * HS + col;
                    float v[8];
                    gather8(red, ksplit, HS, row, col, P + lw.proj_b + gcol, v);
                    add8(v, *reinterpret_cast<const uint4*>(bufN + row * pe + gcol * 2));
                    broadcast16(X + row * pe + gcol * 2, pack8(v));
                }
            }
            MG_PROF(5)
            cluster_sync_all();                // B: x2 is in X everywhere
            MG_PROF(6)
            // ---- P5: m = ln_2(x2);  P6: gelu(c_fc(m)) for this CTA's columns, all-gathered into bufG ----
            layernorm_rows(X, bufN, pe, E, P + lw.ln2_g, P + lw.ln2_b, a.eps, a.use_ln != 0, warp, lane);
            __syncthreads();
            {
                const int ntiles = FS / 8;
                mma_units(bufN, pe, E, S + lw.fc_w, ntiles, 1, [&](int nt) { return crank * FS + nt * 8; }, F - 1, red, warp,
                          lane);
                __syncthreads();
                const int cpr = FS / 8;
                for (int it = tid; it < MG_ROWS * cpr; it += MG_THREADS) {
                    const int row = it / cpr, col = (it % cpr) * 8, gcol = crank * FS + col;
                    float v[8];
                    gather8(red, 1, FS, row, col, P + lw.fc_b + gcol, v);
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = gelu_tanh(v[e]);
                    broadcast16(bufG + row * pf + gcol * 2, pack8(v));
                }
            }
            MG_PROF(7)
            cluster_sync_all();                // C: gelu output is in bufG everywhere
            MG_PROF(8)
            // ---- P7: out = x2 + c_proj(gelu) for this CTA's columns, all-gathered into Y ----
            {
                const int ntiles = HS / 8;
                const int ksplit = max(1, min(MG_WARPS / ntiles, F / 64));
                mma_units(bufG, pf, F, S + lw.proj2_w, ntiles, ksplit, [&](int nt) { return crank * HS + nt * 8; }, E - 1,
                          red, warp, lane);
                __syncthreads();
                const int cpr = HS / 8;
                for (int it = tid; it < MG_ROWS * cpr; it += MG_THREADS) {
                    const int row = it / cpr, col = (it % cpr) * 8, gcol = crank * HS + col;
                    float v[8];
                    gather8(red, ksplit, HS, row, col, P + lw.proj2_b + gcol, v);
                    add8(v, *reinterpret_cast<const uint4*>(X + row * pe + gcol * 2));
                    broadcast16(Y + row * pe + gcol * 2, pack8(v));
                }
            }
            MG_PROF(9)
            cluster_sync_all();                // D: the block output is in Y everywhere
            MG_PROF(10)
            uint8_t* t = X; X = Y; Y = t;
        }

        // ---- ln_f, tied logits for this CTA's vocabulary rows -> Z of CTA 0 ----
        layernorm_rows(X, bufN, pe, E, P + a.lnf_g, P + a.lnf_b, a.eps, a.use_ln != 0, warp, lane);
        __syncthreads();
        {
            const int ntiles = VS / 8;
            const int ksplit = max(1, min(MG_WARPS / ntiles, E / 64));
            mma_units(bufN, pe, E, S + a.wte_sh, ntiles, ksplit, [&](int nt) { return crank * VS + nt * 8; }, V - 1, red,
                      warp, lane);
            __syncthreads();
            const int cpr = VS / 8;
            const uint32_t zbase = map_to_cta(smem_u32(Z), 0);
            for (int it = tid; it < MG_ROWS * cpr; it += MG_THREADS) {
                const int row = it / cpr, col = (it % cpr) * 8, gcol = crank * VS + col;
                float v[8];
                gather8(red, ksplit, VS, row, col, nullptr, v);
                const uint32_t dst = zbase + (row * sm.zp + gcol) * 4;
                st_cluster_v4(dst, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])));
                st_cluster_v4(dst + 16, make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7])));
            }
        }
        MG_PROF(11)
        cluster_sync_all();                    // E: all logits are in CTA 0
        MG_PROF(12)
        if (crank == 0 && warp < G) {
            const int b = s0 + warp;
            float u;
            int chosen = sample_row(Z + warp * sm.zp, V, a.inv_temperature, a.greedy, a.seed_lo, a.seed_hi,
                                    static_cast<uint32_t>(a.seq_base + b), static_cast<uint32_t>(step), lane, &u);
            if (a.forced != nullptr) {
                const int f = a.forced[static_cast<size_t>(b) * a.steps + step];
                if (f >= 0) chosen = f;
            }
            if (lane == 0) {
                a.out_ids[static_cast<size_t>(b) * a.steps + step] = chosen;
                if (a.uniforms != nullptr) a.uniforms[static_cast<size_t>(b) * a.steps + step] = u;
            }
            if (lane < MG_CL) st_cluster_u32(map_to_cta(smem_u32(&toks[warp]), lane), static_cast<uint32_t>(chosen));
            if (a.logits_out != nullptr && step == a.steps - 1)
                for (int c = lane; c < V; c += 32) a.logits_out[static_cast<size_t>(b) * V + c] = Z[warp * sm.zp + c];
        }
        MG_PROF(13)
        cluster_sync_all();                    // F: next tokens are everywhere; Z (= ring of CTA 0) is free again
        MG_PROF(14)
    }
    if (profiling)
        for (int i = 0; i < 16; ++i) a.prof[i] = prof_acc[i];
#undef MG_PROF
}

static MegaSmem mega_smem_layout(int E, int F, int V, int nst) {
    MegaSmem s{};
    const int HS = E / MG_CL, FS = F / MG_CL, VS = 8 * ((V + 63) / 64);
    s.pe = 2 * E + 64;
    s.pf = 2 * F + 64;
    int off = 0;
    auto take = [&](int bytes) { int o = off; off = (off + bytes + 127) & ~127; return o; };
    s.buf0 = take(MG_ROWS * s.pe);
    s.buf1 = take(MG_ROWS * s.pe);
    s.bufn = take(MG_ROWS * s.pe);
    s.bufg = take(MG_ROWS * s.pf);
    s.qkv = take(MG_ROWS * 3 * HS * 2);
    // partial sums: the widest phase is c_fc (FS columns) or a K-split phase (<= 16 / ntiles splits of HS or VS columns)
    int red_cols = FS > 3 * HS ? FS : 3 * HS;
    {
        const int nt = HS / 8, ks = nt >= MG_WARPS ? 1 : MG_WARPS / nt;
        if (ks * HS > red_cols) red_cols = ks * HS;
        const int ntv = VS / 8, ksv = ntv >= MG_WARPS ? 1 : MG_WARPS / ntv;
        if (ksv * VS > red_cols) red_cols = ksv * VS;
    }
    s.red = take(MG_ROWS * red_cols * 4);
    s.nst = nst;
    s.zp = MG_CL * VS;
    int ring_bytes = MG_WARPS * nst * 2 * MG_CHUNK;
    if (MG_ROWS * s.zp * 4 > ring_bytes) ring_bytes = MG_ROWS * s.zp * 4;
    s.ring = take(ring_bytes);
    s.bars = take(MG_WARPS * nst * 8);
    s.toks = take(MG_ROWS * 4);
    s.total = off;
    return s;
}

template <int D>
static int launch_mega(const MegaArgs& args, const MegaSmem& sm, int max_clusters_hint, cudaStream_t s) {
    auto kernel = decode_mega_kernel<D>;
    static int configured_smem = 0;
    if (configured_smem < sm.total) {
        CB200_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sm.total));
        configured_smem = sm.total;
    }
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = MG_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(MG_THREADS);
    cfg.dynamicSmemBytes = sm.total;
    cfg.stream = s;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.gridDim = dim3(MG_CL);
    int resident = 0;
    CB200_CUDA_OK(cudaOccupancyMaxActiveClusters(&resident, kernel, &cfg));
    CB200_REQUIRE(resident >= 1, "the cluster decode kernel does not fit on this device");
    if (max_clusters_hint > 0 && max_clusters_hint < resident) resident = max_clusters_hint;
    // one wave when the batch fits (<= 16 sequences per cluster), otherwise 16-sequence clusters in several waves
    int ncl = args.B < resident ? args.B : resident;
    if (static_cast<long long>(ncl) * MG_ROWS < args.B) ncl = (args.B + MG_ROWS - 1) / MG_ROWS;
    cfg.gridDim = dim3(ncl * MG_CL);
    CB200_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, args, sm));
    note_launch(1);
    return 0;
}

template <int D>
static int mega_capacity(const MegaSmem& sm) {
    auto kernel = decode_mega_kernel<D>;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sm.total) != cudaSuccess) return -1;
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = MG_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(MG_THREADS);
    cfg.gridDim = dim3(MG_CL);
    cfg.dynamicSmemBytes = sm.total;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int resident = 0;
    if (cudaOccupancyMaxActiveClusters(&resident, kernel, &cfg) != cudaSuccess) return -1;
    return resident;
}

// Number of 8-CTA clusters of the persistent decode kernel that are co-resident on the current device.
int decode_mega_capacity(int E, int V, int D) {
    int nst = 4;
    MegaSmem sm = mega_smem_layout(E, 4 * E, V, nst);
    while (sm.total > 227 * 1024 && nst > 2) sm = mega_smem_layout(E, 4 * E, V, --nst);
    switch (D) {
        case 16: return mega_capacity<16>(sm);
        case 32: return mega_capacity<32>(sm);
        default: return mega_capacity<64>(sm);
    }
}

bool decode_mega_supported(int E, int H, int D, int V, int L) {
    if (!(E == 256 || E == 512)) return false;
    if (H % MG_CL != 0 || H * D != E) return false;
    if (!(D == 16 || D == 32 || D == 64)) return false;
    if (L > MG_MAX_LAYERS || V < 1 || V > 4096) return false;
    const MegaSmem sm = mega_smem_layout(E, 4 * E, V, 2);
    return sm.total <= 227 * 1024;
}

int decode_mega(const MegaArgs& args, int D, int max_clusters, cudaStream_t s) {
    CB200_REQUIRE(decode_mega_supported(args.E, args.H, D, args.V, args.L), "shape not supported by the cluster decode kernel");
    if (args.B == 0 || args.steps == 0) return 0;
    int nst = 4;
    MegaSmem sm = mega_smem_layout(args.E, args.F, args.V, nst);
    while (sm.total > 227 * 1024 && nst > 2) sm = mega_smem_layout(args.E, args.F, args.V, --nst);
    switch (D) {
        case 16: return launch_mega<16>(args, sm, max_clusters, s);
        case 32: return launch_mega<32>(args, sm, max_clusters, s);
        default: return launch_mega<64>(args, sm, max_clusters, s);
    }
}

}  // namespace cb200
