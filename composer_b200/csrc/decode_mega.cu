// Persistent cluster decode: the whole generation loop (all steps, all layers)
// in ONE kernel launch.
//
// Sequences never interact while decoding, so the batch is split over thread
// block clusters of 4 or 8 CTAs, each cluster owning up to 8 sequences for the
// whole generation, and there is no grid-wide synchronisation at all: the only
// barriers are hardware cluster barriers.  Inside a cluster every linear layer
// is split over the CTAs by output columns; c_attn is split by heads, so a CTA
// computes q, k, v of its own heads, appends k, v to the cache and attends for
// those heads without any exchange.  The activations that the next layer needs
// in full (attention output, x2, gelu output, block output: <= 8 rows each) are
// all-gathered by writing the CTA's column slice into the shared memory of
// every peer (DSMEM), 4 cluster barriers per decoder block.
//
//  * Linear layers run transposed, out^T = W x^T: the 16 rows of an m16n8k16
//    MMA are 16 output columns, its 8 columns are the 8 sequences.  A pack
//    kernel re-lays the bf16 weights once per generation into the order in
//    which each CTA consumes them: a stream of 4 KB slots (16 output columns x
//    128 of K in A-fragment order, one 16-byte shared-memory read per MMA, +
//    the 16 biases).
//  * KV cache, [L, B, H, t_max, 2, d_h] (k and v of a token adjacent): a warp
//    owns one (sequence, head) pair at a time and streams it in 4 KB stages;
//    scores and P.V run on the tensor cores (scores as K q: the 16 tokens of a
//    tile are the MMA rows, the query is one column of the B operand; the
//    probabilities travel to the A-operand layout of P.V with two shuffles per
//    tile; V rows through ldmatrix.trans, in the order the packed pairs arrive
//    in), the softmax stays online.  With few (sequence, head) pairs per CTA
//    (small batches) a pair's chunks are dealt to 2-4 warps whose partial
//    softmax states are merged through shared memory.
//  * Everything a warp reads from global memory on the hot path, weight slots
//    and KV stages alike, is one static sequence of copy jobs in the warp's own
//    program order (it depends on the step, not on data).  The warp owns 2-3
//    ring stages; after consuming job k it issues job k + stages (one TMA bulk
//    copy of 4 KB by one lane, completion on an mbarrier), so the weights of a
//    GEMM phase are requested while the previous phase or the attention is
//    still running.  L2 eviction hints keep the weight stream resident.  The
//    16-byte pieces of a cached record are XOR-swizzled (in global memory, so
//    that a chunk stays one linear copy): ldmatrix on the stage is then free of
//    bank conflicts, which were what bounded the attention phase before.
// The first CTA of each cluster draws the tokens (same Philox / inverse-CDF
// rule as logits_sample_kernel) and broadcasts them.
//
// Reference semantics: Transformer.call with `past=` (composer/models/
// transformer.py:423-437, 583-597, 735-833) and the sampling rule of
// cli.py:663-676.  Rounding points (bf16 LayerNorm outputs, q/k/v, attention
// output, residual stream, gelu output) are those of the per-step kernels in
// decode.cu, which stay as the path for shapes this kernel does not cover.
#include "decode.h"
#include "decode_common.cuh"
#include "mma_sync.cuh"

#include <cstdlib>
#include <vector>

namespace cb200 {

constexpr int MG_WARPS = 16;
constexpr int MG_THREADS = MG_WARPS * 32;
constexpr int MG_ROWS = 8;          // sequences per cluster: rows 0 .. 7 of the m16n8k16 MMAs (rows 8 .. 15 are fed zeros)
// tokens per KV ring stage (k|v records of one head): 4 KB whatever the head size
#define MG_CT(D) ((D) == 64 ? 16 : (D) == 32 ? 32 : 64)
constexpr int MG_SLOT_W = 4096;            // weight bytes of a slot: 16 output columns x 128 of K in A-fragment order
constexpr int MG_SLOT = MG_SLOT_W + 64;    // + the 16 biases of those columns

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float warp_max_f32(float v) {        // sm_100a: warp-wide fp32 maximum in one instruction
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v2(uint32_t addr, uint32_t x, uint32_t y) {
    asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// A remote store that reports its bytes to an mbarrier of the destination CTA (STAS): the receiver learns that an
// all-gather is complete by waiting for the byte count on its own mbarrier, and nobody executes a fence.
// (`barrier.cluster.arrive.release` / `wait.acquire` compile to MEMBAR.ALL.GPU + ERRBAR in front of the arrive and
// CCTL.IVALL after the wait, in every thread: cluster scope is implemented at GPU scope.  Plain remote stores
// followed by one `mbarrier.arrive.release.cluster` per sending warp were measured slower than the cluster barrier:
// the MEMBAR of a single warp costs as much as that of all warps together.)
__device__ __forceinline__ void st_async_v4(uint32_t addr, const uint4& v, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w), "r"(bar)
                 : "memory");
}
// L2 eviction policies of the copies: the KV cache is streamed once per step (evict first), the weight stream is
// re-read by every cluster every step (evict last), so that 1 GB of cache reads per step does not push the 13 MB of
// weights out of L2.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// TMA bulk copy global -> shared (one instruction, one thread), completion as transaction bytes on an mbarrier.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct MegaSmem {        // shared-memory layout and weight-stream plan (computed on the host, identical in every CTA)
    int pe, pf;          // row pitch of the [8, E] and [8, F] bf16 buffers (2E + 64, 2F + 64: conflict-free operand reads)
    int buf0, buf1, bufn, bufg, qkv, red, zbuf, jobtab, ring, bars, toks, prof, total;
    int nst;             // ring stages per warp ...
    int nst_extra;       // ... plus one more for the first nst_extra warps (whatever still fits)
    int zp;              // floats per row of the logits matrix (CTA 0)
    int vs;              // vocabulary rows per CTA (whole 16-row tiles)
    int ks_attn, ks_proj, ks_fc, ks_proj2, ks_logits;        // K split of each phase
    int n_attn, n_proj, n_fc, n_proj2, n_logits;             // slots per phase
    int per_layer, per_step;                                 // slots per decoder block / per step
};

// One linear layer as this CTA sees it, computed transposed (out^T = W x^T) so that the 16 rows of an m16n8k16 MMA
// are 16 output columns and its 8 columns are the <= 8 sequences of the cluster: `ntiles` tiles of 16 output columns,
// K split `ksplit` ways; unit u = (tile u % ntiles, K part u / ntiles) belongs to warp u % 16 and occupies `sub`
// consecutive slots (128 of K each) of the weight stream.
struct Phase {
    int ntiles, ksplit, sub;
    int nt0, ks0;        // this warp's unit when the K is split (unit = warp): tile and K part
};

// ---------------------------------------------------------------------------
// Per-warp job ring.  Everything a warp reads from global memory on the hot path, the weight slots of its GEMM
// units and the KV stages of its (sequence, head) pairs, is one sequence of TMA bulk copies in the warp's own
// program order, which is fully static (it depends on the step, not on data).  The warp owns NST stages of
// MG_STAGE bytes; after consuming job k it issues job k + NST into the stage it has just freed, so its next
// NST jobs are always in flight, across phase and barrier boundaries: the weights of a GEMM phase were requested
// while the previous phase (or the attention) was still running.
// ---------------------------------------------------------------------------
constexpr int MG_STAGE = MG_SLOT;          // >= the 4 KB of a KV stage

// Byte offset of 16-byte piece `c` of token `t` inside a KV stage.  A k|v record is 4 D bytes (4, 8 or 16 pieces);
// ldmatrix reads one piece of 8 tokens, which would land in 2 (d_h 16) or 1 (d_h >= 32) of the 8 16-byte bank
// groups, a 4- or 8-way conflict that made shared memory the limiter of the attention phase.  The piece position
// is XOR-ed with token bits so that both row sets the attention reads cover all 8 groups: 8 consecutive tokens (K,
// the score MMA) and tokens {j, j + 8 : j even} or {j, j + 8 : j odd} of a 16-token tile (V, in the order the
// probabilities come out of the softmax).  d_h 16: two records share a 128-byte line and the XOR runs over the 8
// positions of the line.  The appended records are written to the cache in this form.
template <int D>
__device__ __forceinline__ uint32_t kv_piece_off(int t, int c) {
    constexpr int REC = 4 * D;
    if (D == 16) {
        const int pos = (((t & 1) << 2) | c) ^ (((t >> 1) & 3) | (((t >> 3) & 1) << 2));
        return static_cast<uint32_t>((t >> 1) * (2 * REC) + (pos << 4));
    }
    const int sw = (t ^ ((t >> 3) & 1)) & 7;       // (bit 3 only: the same for every 16-token tile of the stage)
    return static_cast<uint32_t>(t * REC + ((c ^ sw) << 4));
}

struct JobPlan {           // what the cursor needs to enumerate this warp's jobs (uniform per warp)
    const uint8_t* wsrc;   // this CTA's weight stream (one step; it repeats every step)
    const __nv_bfloat16* cache;
    long long layer_stride;
    uint64_t w_policy, kv_policy;
    const uint16_t* tab;   // this warp's weight jobs of one decoder block: slots before the attention (c_attn), slots
                           // after it (c_proj, c_fc, mlp c_proj), then the logits slots; relative stream positions
    int n_pre, n_post, n_logit;
    int warp, steps, L, s0, nmine, hpc_shift, HPC, H, crank, t_max, per_layer;
    int ss, split;         // a pair's chunks are dealt to 1 << ss adjacent warps; this warp takes chunks = split mod 2^ss
};
constexpr int MG_TAB = 32;                 // table entries per warp: [0, 8) pre, [8, 24) post, [24, 32) logits

struct JobCursor {
    int step, l, k, pi;                    // k: weight jobs of this block issued so far; pi: next KV item
    int kv_left;                           // chunks of the current KV item still to issue, the next one at kv_ptr
    const uint8_t* kv_ptr;
};

struct JobRing {
    uint8_t* base;         // this warp's stages
    uint64_t* bars;        // this warp's mbarriers
    int stage;             // stage of the next job to consume ...
    uint32_t phase;        // ... and the parity of its mbarrier phase
    int nst;
    JobCursor cur;         // next job to issue (always `nst` jobs ahead of `count`)
    long long* wait_prof;  // diagnostic accumulator or null
};

__device__ __forceinline__ int units_of(int warp, int n) { return n > warp ? (n - warp + MG_WARPS - 1) / MG_WARPS : 0; }

// Issues the job at the cursor into `stage` and advances the cursor (all lanes advance it, lane 0 copies).  D selects
// the KV geometry.
template <int D>
__device__ __forceinline__ void job_issue(const JobPlan& p, JobCursor& c, uint8_t* stage, uint64_t* bar, int lane) {
    constexpr int CT = MG_CT(D);
    constexpr uint32_t CHUNK_BYTES = CT * 4 * D;
    const uint8_t* src;
    bool kv = true;
    if (c.kv_left > 0) {                            // the common case: the next chunk of the current (sequence, head) item
        src = c.kv_ptr;
        --c.kv_left;
    } else {
        for (;;) {
            if (c.step >= p.steps) return;
            if (c.l < p.L) {
                if (c.k < p.n_pre) { src = p.wsrc + static_cast<size_t>(c.l * p.per_layer + p.tab[c.k]) * MG_SLOT; ++c.k; kv = false; break; }
                if (c.pi < p.nmine) {
                    // chunks of an item of this warp in this step: split, split + S, ... (the same for all its items)
                    const int nchunks = (c.step + CT - 1) / CT;
                    const int nck = nchunks > p.split ? (nchunks - p.split + (1 << p.ss) - 1) >> p.ss : 0;
                    if (nck > 0) {
                        const int q = (p.warp + MG_WARPS * c.pi) >> p.ss;
                        const int b = p.s0 + (q >> p.hpc_shift), h = p.crank * p.HPC + (q & (p.HPC - 1));
                        const size_t off = ((static_cast<size_t>(b) * p.H + h) * p.t_max + static_cast<size_t>(p.split) * CT) * (2 * D);
                        src = reinterpret_cast<const uint8_t*>(p.cache + static_cast<size_t>(c.l) * p.layer_stride + off);
                        ++c.pi;
                        c.kv_left = nck - 1;
                        break;
                    }
                    c.pi = p.nmine;
                }
                const int kk = c.k - p.n_pre;
                if (kk < p.n_post) { src = p.wsrc + static_cast<size_t>(c.l * p.per_layer + p.tab[8 + kk]) * MG_SLOT; ++c.k; kv = false; break; }
                ++c.l; c.k = 0; c.pi = 0;
            } else {
                if (c.k < p.n_logit) { src = p.wsrc + static_cast<size_t>(p.L * p.per_layer + p.tab[24 + c.k]) * MG_SLOT; ++c.k; kv = false; break; }
                ++c.step; c.l = 0; c.k = 0; c.pi = 0;
            }
        }
    }
    // one TMA bulk copy per job, issued by lane 0: a weight slot, or a whole chunk of a head's k|v records (the cache
    // is zero-filled before the generation, so the tokens of the chunk that are not cached yet read as zeros)
    if (kv) c.kv_ptr = src + (static_cast<size_t>(CHUNK_BYTES) << p.ss);
    const uint32_t bytes = kv ? CHUNK_BYTES : static_cast<uint32_t>(MG_SLOT);
    if (lane == 0) {
        mbar_expect_tx(bar, bytes);
        bulk_g2s(smem_u32(stage), src, bytes, bar, kv ? p.kv_policy : p.w_policy);
    }
}

// L2 prefetch of this warp's KV jobs [lo, hi) of the attention phase of (`step`, `layer`), one job per lane: the
// attention phase streams the cache at HBM speed while the GEMM phases leave HBM idle, so the GEMM phases can ask L2
// for the first chunks of the coming attention phase.  OFF by default (CB200_DECODE_KV_PREFETCH_MB): measured, the
// hits save less than the prefetch traffic costs the L2-resident weight slots the GEMM phases wait for.
template <int D>
__device__ __forceinline__ void kv_prefetch_l2(const JobPlan& p, int layer, int step, int lo, int hi, int lane) {
    constexpr int CT = MG_CT(D);
    constexpr int CHUNK_BYTES = CT * 4 * D;
    const int nchunks = (step + CT - 1) / CT;
    const int nck = nchunks > p.split ? (nchunks - p.split + (1 << p.ss) - 1) >> p.ss : 0;    // chunks per item of this warp
    int n = lo + lane;
    if (nck == 0 || n >= hi) return;
    int pi = 0;
    while (pi < p.nmine && n >= nck) { n -= nck; ++pi; }
    if (pi >= p.nmine) return;
    const int q = (p.warp + MG_WARPS * pi) >> p.ss;
    const int b = p.s0 + (q >> p.hpc_shift), h = p.crank * p.HPC + (q & (p.HPC - 1));
    const size_t off = ((static_cast<size_t>(b) * p.H + h) * p.t_max + static_cast<size_t>(p.split + (n << p.ss)) * CT) * (2 * D);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.cache + static_cast<size_t>(layer) * p.layer_stride + off), "r"(CHUNK_BYTES)
                 : "memory");
}

// Waits for the warp's next job and returns its stage.  (The stage index and the mbarrier phase are tracked
// incrementally: an integer division by the runtime stage count costs more than the MMAs of a slot.)
template <bool PROF>
__device__ __forceinline__ uint8_t* ring_acquire(const JobRing& r) {
    if (PROF && r.wait_prof != nullptr) {          // diagnostic: cycles this thread waits for its jobs
        const long long t0 = clock64();
        mbar_wait(&r.bars[r.stage], r.phase);
        *r.wait_prof += clock64() - t0;
    } else {
        mbar_wait(&r.bars[r.stage], r.phase);
    }
    return r.base + r.stage * MG_STAGE;
}

// The warp is done with the stage of its current job: it is re-armed with the job NST ahead.
template <int D>
__device__ __forceinline__ void ring_release(JobRing& r, const JobPlan& p, int lane) {
    __syncwarp();
    job_issue<D>(p, r.cur, r.base + r.stage * MG_STAGE, &r.bars[r.stage], lane);
    if (++r.stage == r.nst) { r.stage = 0; r.phase ^= 1u; }
}

// acc (16 output columns x 8 sequences) += slot (16 x 128 weights, A operand, one 16-byte read per MMA) times
// x^T: b0 points at (sequence g, first k of the slot + 8 * tig) of the activations (B operand: the k order inside
// a 32-wide block is permuted identically on both sides).
__device__ __forceinline__ void slot_mma(float (&acc)[4], const uint8_t* slot, const uint8_t* b0, int lane) {
    // four independent accumulation chains (one per 32-wide k-block): a dependent mma.sync chain of 8 would expose
    // the tensor-core latency 8 times per slot
    float part[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint4 x = *reinterpret_cast<const uint4*>(b0 + i * 64);
        const uint4 wa = *reinterpret_cast<const uint4*>(slot + (2 * i * 32 + lane) * 16);
        const uint4 wb = *reinterpret_cast<const uint4*>(slot + ((2 * i + 1) * 32 + lane) * 16);
        part[i][0] = part[i][1] = part[i][2] = part[i][3] = 0.f;
        mma_16816(part[i], wa.x, wa.y, wa.z, wa.w, x.x, x.y);
        mma_16816(part[i], wb.x, wb.y, wb.z, wb.w, x.z, x.w);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] += (part[0][e] + part[1][e]) + (part[2][e] + part[3][e]);
}

// Diagnostic (PROF instantiation only): finer marks inside a linear phase, cycles of one thread since the last mark.
struct FineProf {
    long long* acc;        // 6 counters of the current phase kind, or null
    long long t;
    __device__ __forceinline__ void mark(int i) {
        if (acc != nullptr) { const long long now = clock64(); acc[i] += now - t; t = now; }
    }
};

// Runs this warp's units of `ph` on the activations X[8, K] (bf16 in shared memory, row pitch `pitch` bytes); the
// weight slots are the warp's next jobs.  K-split partials meet in `red`.  The warp that owns K part 0 of a tile
// calls epi(col, seq, lo, hi) once per lane with the complete sums (bias added) of sequence `seq` at output columns
// (col, col + 1) = lo and (col + 8, col + 9) = hi, col relative to this CTA's slice.  All 16 warps must call this
// (it may contain a __syncthreads).  There is one call site (the phase loop of the kernel): the kernel's code has
// to stay within the instruction cache, a phase is only a few hundred instructions long.
template <int D, bool PROF, bool DEFER, typename Epi>
__device__ __forceinline__ void run_phase(const Phase& ph, const uint8_t* X, int pitch, JobRing& ring, const JobPlan& plan,
                                          float* red, int warp, int lane, FineProf fine, Epi epi) {
    const int g = lane >> 2, tig = lane & 3;
    if (PROF) fine.mark(0);
    const uint8_t* b0 = X + g * pitch + tig * 16;
    const int nunits = ph.ntiles * ph.ksplit;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float bias_lo = 0.f, bias_hi = 0.f;
    int nt = 0;
    bool owner = false;                            // this warp holds K part 0 of tile nt in acc
    bool held = false;                             // the stage of the last slot has not been released yet
    for (int u = warp; u < nunits; u += MG_WARPS) {
        // with several units per warp (then ksplit == 1) the previous tile is finished first
        if (owner) {
            acc[0] += bias_lo; acc[1] += bias_lo; acc[2] += bias_hi; acc[3] += bias_hi;
            const bool odd = g & 1;
            const float r0 = __shfl_xor_sync(0xffffffffu, odd ? acc[0] : acc[1], 4);
            const float r1 = __shfl_xor_sync(0xffffffffu, odd ? acc[2] : acc[3], 4);
            epi(nt * 16 + (g & ~1), 2 * tig + (odd ? 1 : 0), odd ? make_float2(r0, acc[1]) : make_float2(acc[0], r0),
                odd ? make_float2(r1, acc[3]) : make_float2(acc[2], r1));
            acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
        }
        nt = (ph.ksplit == 1) ? u : ph.nt0;          // (no division here: it would cost as much as the MMAs of a slot)
        const int ks = (ph.ksplit == 1) ? 0 : ph.ks0;
        for (int j = 0; j < ph.sub; ++j) {
            const uint8_t* slot = ring_acquire<PROF>(ring);
            if (PROF) fine.mark(1);
            if (j == 0) {
                bias_lo = *reinterpret_cast<const float*>(slot + MG_SLOT_W + g * 4);
                bias_hi = *reinterpret_cast<const float*>(slot + MG_SLOT_W + (g + 8) * 4);
            }
            slot_mma(acc, slot, b0 + (ks * ph.sub + j) * 256, lane);
            if (PROF) fine.mark(2);
            held = DEFER && (j + 1 == ph.sub) && (u + MG_WARPS >= nunits);
            if (!held) {
                ring_release<D>(ring, plan, lane);
                if (PROF) fine.mark(3);
            }
        }
        owner = ks == 0;
        if (ks > 0) *reinterpret_cast<float4*>(red + (((ks - 1) * ph.ntiles + nt) * 32 + lane) * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
    if (ph.ksplit > 1) {
        __syncthreads();
        if (PROF) fine.mark(4);
        if (owner)
            for (int k2 = 1; k2 < ph.ksplit; ++k2) {
                const float4 p = *reinterpret_cast<const float4*>(red + (((k2 - 1) * ph.ntiles + nt) * 32 + lane) * 4);
                acc[0] += p.x; acc[1] += p.y; acc[2] += p.z; acc[3] += p.w;
            }
    }
    if (owner) {
        acc[0] += bias_lo; acc[1] += bias_lo; acc[2] += bias_hi; acc[3] += bias_hi;
        // lanes g and g ^ 1 trade one sequence each: afterwards a lane holds two adjacent columns of one sequence
        const bool odd = g & 1;
        const float r0 = __shfl_xor_sync(0xffffffffu, odd ? acc[0] : acc[1], 4);
        const float r1 = __shfl_xor_sync(0xffffffffu, odd ? acc[2] : acc[3], 4);
        epi(nt * 16 + (g & ~1), 2 * tig + (odd ? 1 : 0), odd ? make_float2(r0, acc[1]) : make_float2(acc[0], r0),
            odd ? make_float2(r1, acc[3]) : make_float2(acc[2], r1));
    }
    if (PROF) fine.mark(5);
    // DEFER (few sequences per cluster, every phase is a latency chain): the stage of the warp's last slot is re-armed
    // only now; finding and issuing the next job is ~100 scalar instructions, which otherwise sit between the MMAs
    // and the K-split barrier / the epilogue.  (With 8 sequences per cluster the attention phase is paced by the
    // ring, and issuing later costs more than it saves: 250 vs 239 us per step at 256 sequences.)
    if (held) {
        ring_release<D>(ring, plan, lane);
        if (PROF) fine.mark(3);
    }
}

__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

// Writes one bf16 pair to the same shared-memory offset of all CL CTAs of the cluster (rows without a sequence
// are skipped: DSMEM bandwidth is the cost of the all-gathers).
template <int CL>
__device__ __forceinline__ void broadcast_u32(const uint8_t* local, uint32_t val, bool valid) {
    if (!valid) return;
    const uint32_t a = smem_u32(local);
#pragma unroll
    for (int r = 0; r < CL; ++r) st_cluster_u32(map_to_cta(a, r), val);
}

// A 16-byte piece of this CTA's copy of an activation to the same place in every CTA of the cluster (this one
// included: its mbarrier counts those bytes too), each store counted on the mbarrier `bar` (a shared::cta address,
// the same in every CTA) of its destination.  The window of a CTA in the cluster's shared address space keeps the
// offsets, so one `mapa` per destination serves the barrier and the data.
template <int CL>
__device__ __forceinline__ void broadcast16_async(const uint8_t* local, uint32_t bar) {
    const uint4 v = *reinterpret_cast<const uint4*>(local);
    const uint32_t d = smem_u32(local) - bar;
#pragma unroll
    for (int r = 0; r < CL; ++r) {
        const uint32_t rb = map_to_cta(bar, r);
        st_async_v4(rb + d, v, rb);
    }
}

// gamma / beta of one LayerNorm as the lanes of a warp need them (lane = 16-byte chunk of the row, two chunks for
// E = 512); fetched a phase ahead so that the L2 round trip is hidden.
struct LnFrag {
    float4 g[2][2], b[2][2];
};

__device__ __forceinline__ void ln_prefetch(LnFrag& f, const float* gamma, const float* beta, int E, int lane) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int ch = lane + 32 * i;
        if (ch < E / 8) {
            f.g[i][0] = __ldg(reinterpret_cast<const float4*>(gamma + ch * 8)); f.g[i][1] = __ldg(reinterpret_cast<const float4*>(gamma + ch * 8 + 4));
            f.b[i][0] = __ldg(reinterpret_cast<const float4*>(beta + ch * 8)); f.b[i][1] = __ldg(reinterpret_cast<const float4*>(beta + ch * 8 + 4));
        }
    }
}

// LayerNorm of the 16 rows of `src` into `dst` (warp = row), both [16, E] bf16 with pitch pe.  Two passes in
// registers like layernorm_fwd_kernel; the output is rounded to bf16 (what the next GEMM consumes).
__device__ __forceinline__ void layernorm_rows(const uint8_t* src, uint8_t* dst, int pe, int E, const LnFrag& f, float eps,
                                               bool enabled, int rows, int warp, int lane) {
    if (warp >= rows) return;                 // (rows without a sequence keep the zeros of the start: their MMA columns are not used)
    const uint8_t* s = src + warp * pe;
    uint8_t* d = dst + warp * pe;
    const int nchunk = E / 8;                 // 16-byte chunks per row: 32 (E 256) or 64 (E 512)
    if (!enabled) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int ch = lane + 32 * i;
            if (ch < nchunk) *reinterpret_cast<uint4*>(d + ch * 16) = *reinterpret_cast<const uint4*>(s + ch * 16);
        }
        return;
    }
    // sum and sum of squares in one pass and one butterfly (fp32; the inputs are bf16 values of a few units, the
    // cancellation error is far below the bf16 rounding of the output): half the dependent shuffle rounds
    float v[2][8];
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int ch = lane + 32 * i;
        if (ch < nchunk) {
            const uint4 raw = *reinterpret_cast<const uint4*>(s + ch * 16);
            const float2 a = unpack_bf16(raw.x), b = unpack_bf16(raw.y), c = unpack_bf16(raw.z), e = unpack_bf16(raw.w);
            v[i][0] = a.x; v[i][1] = a.y; v[i][2] = b.x; v[i][3] = b.y; v[i][4] = c.x; v[i][5] = c.y; v[i][6] = e.x; v[i][7] = e.y;
            sum += ((v[i][0] + v[i][1]) + (v[i][2] + v[i][3])) + ((v[i][4] + v[i][5]) + (v[i][6] + v[i][7]));
            sq += ((v[i][0] * v[i][0] + v[i][1] * v[i][1]) + (v[i][2] * v[i][2] + v[i][3] * v[i][3])) +
                  ((v[i][4] * v[i][4] + v[i][5] * v[i][5]) + (v[i][6] * v[i][6] + v[i][7] * v[i][7]));
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[i][k] = 0.f;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    const float inv_e = 1.0f / E;
    const float mean = sum * inv_e;
    const float rstd = rsqrtf(fmaxf(sq * inv_e - mean * mean, 0.f) + eps);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int ch = lane + 32 * i;
        if (ch < nchunk) {
            float o[8];
            o[0] = (v[i][0] - mean) * rstd * f.g[i][0].x + f.b[i][0].x; o[1] = (v[i][1] - mean) * rstd * f.g[i][0].y + f.b[i][0].y;
            o[2] = (v[i][2] - mean) * rstd * f.g[i][0].z + f.b[i][0].z; o[3] = (v[i][3] - mean) * rstd * f.g[i][0].w + f.b[i][0].w;
            o[4] = (v[i][4] - mean) * rstd * f.g[i][1].x + f.b[i][1].x; o[5] = (v[i][5] - mean) * rstd * f.g[i][1].y + f.b[i][1].y;
            o[6] = (v[i][6] - mean) * rstd * f.g[i][1].z + f.b[i][1].z; o[7] = (v[i][7] - mean) * rstd * f.g[i][1].w + f.b[i][1].w;
            *reinterpret_cast<uint4*>(d + ch * 16) = pack8(o);
        }
    }
}

// sample_row of decode_common.cuh with the exponentials cached in place (z is overwritten): same expressions,
// same order of additions, hence the same token as the per-step sampler.
__device__ __forceinline__ int sample_row_smem(float* z, int V, float inv_temperature, int greedy, uint32_t seed_lo,
                                               uint32_t seed_hi, uint32_t seq_index, uint32_t step, int lane, float* u_used) {
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
        const float c = inv_temperature * 1.4426950408889634f;
        const int per = (V + 31) / 32;
        const int lo = lane * per, hi = min(V, lo + per);
        float mass = 0.f;
        for (int i = lo; i < hi; ++i) {
            const float e = exp2f((z[i] - vmax) * c);
            z[i] = e;
            mass += e;
        }
        float prefix = mass;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float t = __shfl_up_sync(0xffffffffu, prefix, o);
            if (lane >= o) prefix += t;
        }
        const float total = __shfl_sync(0xffffffffu, prefix, 31);
        const Philox4 r = philox4x32_10(step, seq_index, 0x5A17u, 0u, seed_lo, seed_hi);
        u = (r.x >> 8) * (1.0f / 16777216.0f);
        const float target = u * total;
        const float before = prefix - mass;
        const bool mine = (target >= before && target < prefix) || (lane == 31 && target >= prefix);
        int pick = -1;
        if (mine) {
            float run = before;
            pick = max(hi - 1, lo);
            for (int i = lo; i < hi; ++i) {
                run += z[i];
                if (target < run) { pick = i; break; }
            }
            if (pick >= V) pick = V - 1;
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, pick >= 0);
        const int src = ballot ? (__ffs(ballot) - 1) : 0;
        chosen = __shfl_sync(0xffffffffu, pick, src);
        if (chosen < 0) chosen = amax;
    }
    *u_used = u;
    return chosen;
}

// PROF: the phase profile (cb200_set_decode_profile) is compiled in; a separate instantiation, because even never-taken
// diagnostic branches on the job path cost the default kernel issue slots.
template <int D, int CL, bool PROF = false, bool AG = true>
__global__ void __launch_bounds__(MG_THREADS, 1)
decode_mega_kernel(const __grid_constant__ MegaArgs a, const __grid_constant__ MegaSmem sm) {
    constexpr int CH = D / 8;                     // 16-byte chunks per head row
    constexpr int CT = MG_CT(D);                  // tokens per KV ring stage
    constexpr int REC = 4 * D;                    // bytes of one cached token of one head: k row | v row
    static_assert(CT * REC <= MG_STAGE, "a KV stage must fit a ring stage");
    constexpr int NT_S = CT / 8;                  // score n-tiles per stage
    constexpr int NT_O = D / 8;                   // output n-tiles
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int g = lane >> 2, tig = lane & 3;
    const int crank = static_cast<int>(cluster_ctarank());
    const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;
    // sequences of this cluster: the first (B % ncl) clusters take one more
    const int base = a.B / ncl, rem = a.B % ncl;
    const int G = base + (cid < rem ? 1 : 0);
    const int s0 = cid * base + min(cid, rem);
    if (G == 0) return;                            // whole cluster leaves together

    const int E = a.E, H = a.H, V = a.V;
    const int HS = E / CL;                         // columns of the residual stream owned by this CTA (its heads)
    const int FS = a.F / CL;
    const int HPC = H / CL;                        // heads per CTA
    const int VS = sm.vs;
    const int pe = sm.pe, pf = sm.pf;
    uint8_t* bufU = smem + sm.buf0;                // block input x, later x2
    uint8_t* bufW = smem + sm.buf1;                // attention output, later the block output
    uint8_t* bufN = smem + sm.bufn;                // LayerNorm output (ln_1: the residual stream of the block)
    uint8_t* bufG = smem + sm.bufg;                // gelu output [16, F]
    uint8_t* qkvs = smem + sm.qkv;                 // q | k | v of this CTA's heads, [16][3 * HS] bf16
    float* red = reinterpret_cast<float*>(smem + sm.red);
    uint8_t* ring = smem + sm.ring;
    float* Z = reinterpret_cast<float*>(smem + sm.zbuf);   // logits [8][zp], used in CTA 0
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sm.bars);       // [16 warps][NST stages]
    int* toks = reinterpret_cast<int*>(smem + sm.toks);
    const int NST = sm.nst;

    // ---- one-time setup ----
    for (int i = tid * 16; i < sm.bars; i += MG_THREADS * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    if (tid < MG_WARPS * NST + sm.nst_extra + 4) mbar_init(&bars[tid], 1);
    if (tid < MG_ROWS) toks[tid] = (tid < G) ? a.first[s0 + tid] : 0;
    mbar_fence_init();
    __syncthreads();
    // AG: the four all-gathers of a decoder block (attention output, x2, GELU output, block output) complete on
    // mbarriers of the receiving CTA: one arrival (thread 0, which states the bytes of the coming gather) plus
    // the bytes themselves.  A gather can only be sent to a CTA that has passed the previous use of the same barrier
    // and re-armed it: the sender has by then received that CTA's slices of the three gathers in between.
    uint64_t* gbar = bars + MG_WARPS * NST + sm.nst_extra;
    const uint32_t gbar_a = smem_u32(gbar);
    const uint32_t gbytes[4] = {static_cast<uint32_t>(G * a.E * 2), static_cast<uint32_t>(G * a.E * 2),
                                static_cast<uint32_t>(G * a.F * 2), static_cast<uint32_t>(G * a.E * 2)};
    uint32_t gpar = 0;
    if (AG && tid == 0)
        for (int k = 0; k < 4; ++k) mbar_expect_tx(&gbar[k], gbytes[k]);

    // phases of a decoder block and of the head as this CTA sees them (16-column tiles, K split, slots per unit)
    auto make_phase = [&](int ntiles, int ksplit, int K) {
        Phase ph;
        ph.ntiles = ntiles; ph.ksplit = ksplit; ph.sub = K / 128 / ksplit;
        ph.nt0 = warp % ntiles; ph.ks0 = warp / ntiles;
        return ph;
    };
    const Phase ph_attn = make_phase(3 * HS / 16, sm.ks_attn, E);
    const Phase ph_proj = make_phase(HS / 16, sm.ks_proj, E);
    const Phase ph_fc = make_phase(FS / 16, sm.ks_fc, E);
    const Phase ph_proj2 = make_phase(HS / 16, sm.ks_proj2, a.F);
    const Phase ph_logits = make_phase(VS / 16, sm.ks_logits, E);

    // this warp's weight jobs in the order it consumes them (stream positions relative to the block / to the head)
    uint16_t* tab = reinterpret_cast<uint16_t*>(smem + sm.jobtab) + warp * MG_TAB;
    int n_pre = 0, n_post = 0, n_logit = 0;
    {
        auto list = [&](const Phase& ph, int pos0, int first, int& n) {
            for (int u = warp; u < ph.ntiles * ph.ksplit; u += MG_WARPS)
                for (int j = 0; j < ph.sub; ++j) {
                    if (lane == 0) tab[first + n] = static_cast<uint16_t>(pos0 + u * ph.sub + j);
                    ++n;
                }
        };
        list(ph_attn, 0, 0, n_pre);
        list(ph_proj, sm.n_attn, 8, n_post);
        list(ph_fc, sm.n_attn + sm.n_proj, 8, n_post);
        list(ph_proj2, sm.n_attn + sm.n_proj + sm.n_fc, 8, n_post);
        list(ph_logits, 0, 24, n_logit);
        __syncwarp();
    }
    JobPlan plan;
    plan.wsrc = a.wstream + static_cast<size_t>(crank) * sm.per_step * MG_SLOT;
    plan.cache = a.cache; plan.layer_stride = a.layer_stride;
    plan.w_policy = a.l2_hints ? l2_policy_evict_last() : l2_policy_evict_normal();
    plan.kv_policy = a.l2_hints ? l2_policy_evict_first() : l2_policy_evict_normal();
    plan.tab = tab; plan.n_pre = n_pre; plan.n_post = n_post; plan.n_logit = n_logit;
    // With few (sequence, head) pairs in the CTA (small batches) the chunks of a pair are dealt to 2, 4 or 8 adjacent
    // warps, which merge their partial softmax states through shared memory: the attention phase is a serial walk
    // over the context, so its length is what the split shortens.
    int kv_ss = 0;
    while (kv_ss < a.kv_split_log2 && ((G * HPC) << (kv_ss + 1)) <= MG_WARPS) ++kv_ss;
    plan.ss = kv_ss; plan.split = warp & ((1 << kv_ss) - 1);
    plan.warp = warp; plan.steps = a.steps; plan.L = a.L; plan.s0 = s0; plan.nmine = units_of(warp, (G * HPC) << kv_ss);
    plan.hpc_shift = 31 - __clz(HPC); plan.HPC = HPC; plan.H = H; plan.crank = crank; plan.t_max = a.t_max;
    plan.per_layer = sm.per_layer;
    JobRing jr;
    const int first_stage = warp * NST + min(warp, sm.nst_extra);        // stages of the warps before this one
    jr.base = ring + first_stage * MG_STAGE;
    jr.bars = bars + first_stage;
    jr.stage = 0; jr.phase = 0; jr.nst = NST + (warp < sm.nst_extra ? 1 : 0);
    jr.cur = JobCursor{a.step0, 0, 0, 0, 0, nullptr};
    jr.wait_prof = nullptr;
    const bool wait_profiling = PROF && a.prof != nullptr && blockIdx.x == 0 && tid == 0;
#define MG_WAIT_SLOT(k) if (PROF && wait_profiling) jr.wait_prof = prof_acc + 16 + (k);
    for (int st = 0; st < jr.nst; ++st) job_issue<D>(plan, jr.cur, jr.base + st * MG_STAGE, &jr.bars[st], lane);
    cluster_sync_all();                            // every CTA of the cluster is resident before any DSMEM access

    // optional phase profile (cluster 0, CTA 0, thread 0): cycles per phase, accumulated in shared memory (a global
    // read-modify-write per sample would cost more than most of the phases it measures) and written out at the end
    const bool profiling = PROF && a.prof != nullptr && blockIdx.x == 0 && tid == 0;
    long long* prof_acc = reinterpret_cast<long long*>(smem + sm.prof);
    if (profiling)
        for (int i = 0; i < 64; ++i) prof_acc[i] = 0;
    long long prof_t = profiling ? clock64() : 0;
#define MG_PROF(slot)                                                        \
    if (PROF && profiling) {                                                 \
        const long long now_ = clock64();                                    \
        prof_acc[slot] += now_ - prof_t;                                     \
        prof_t = now_;                                                       \
    }
    const float* P = a.params;
    const bool use_ln = a.use_ln != 0;
    LnFrag lnf;
    if (use_ln) ln_prefetch(lnf, P + a.layers[0].ln1_g, P + a.layers[0].ln1_b, E, lane);

    for (int step = a.step0; step < a.steps; ++step) {
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

        uint8_t* X = bufU;                         // block input (full rows)
        uint8_t* Y = bufW;                         // the other full-row buffer
        // One loop over the 4 L + 1 linear phases of a step (c_attn [+ attention], c_proj, c_fc, mlp c_proj per block,
        // then the logits), so that every piece of code exists once: the kernel has to fit the instruction cache.
        for (int p = 0; p <= 4 * a.L; ++p) {
            const int l = p >> 2;
            const int kind = (l == a.L) ? 4 : (p & 3);         // 0 c_attn, 1 c_proj, 2 c_fc, 3 mlp c_proj, 4 logits
            const MegaLayer& lw = a.layers[l < a.L ? l : 0];
            // ---- LayerNorm in front of c_attn (ln_1), c_fc (ln_2) and the logits (ln_f) ----
            if (kind == 0 || kind == 2 || kind == 4) {
                layernorm_rows(X, bufN, pe, E, lnf, a.eps, use_ln, G, warp, lane);
                if (use_ln && kind != 0) {         // (ln_2's parameters are fetched after the attention)
                    const bool more = kind == 2 && l + 1 < a.L;          // next: ln_1 of the next block, else ln_f / ln_1 of block 0
                    const MegaLayer& nx = a.layers[more ? l + 1 : 0];
                    if (kind == 2 && !more) ln_prefetch(lnf, P + a.lnf_g, P + a.lnf_b, E, lane);
                    else ln_prefetch(lnf, P + nx.ln1_g, P + nx.ln1_b, E, lane);
                }
                __syncthreads();
                MG_PROF(kind == 0 ? 1 : 15)
            }
            // ---- the linear layer: this CTA's output columns for all sequences ----
            const Phase ph = kind == 0 ? ph_attn : kind == 1 ? ph_proj : kind == 2 ? ph_fc : kind == 3 ? ph_proj2 : ph_logits;
            const uint8_t* src = (kind == 1) ? Y : (kind == 3) ? bufG : bufN;      // x1 | attention output | ln_2(x2) | gelu
            const int src_pitch = (kind == 3) ? pf : pe;
            // epilogue: where the result goes (all-gathered into every CTA but for q, k, v), what is added to it
            uint8_t* dst = (kind == 1) ? X : (kind == 2) ? bufG : Y;
            const uint8_t* res = (kind == 1) ? bufN : X;                             // residual stream: x1, then x2
            const int dst_pitch = (kind == 2) ? pf : pe;
            const int col0 = crank * ((kind == 2) ? FS : HS);
            const uint32_t zbase = map_to_cta(smem_u32(Z), 0);
            if (a.kv_prefetch > 0 && kind >= 1 && kind <= 3) {          // a third of the budget in front of each GEMM phase
                const bool wrap = l + 1 == a.L;
                if (!wrap || step + 1 < a.steps)
                    kv_prefetch_l2<D>(plan, wrap ? 0 : l + 1, wrap ? step + 1 : step, jr.nst + (kind - 1) * a.kv_prefetch / 3,
                                      jr.nst + kind * a.kv_prefetch / 3, lane);      // (the first jobs of the phase are fetched by the ring itself)
            }
            MG_WAIT_SLOT(kind == 0 ? 0 : kind + 1)
            FineProf fine{nullptr, 0};
            if (PROF && profiling) { fine.acc = prof_acc + 24 + 6 * kind; fine.t = prof_t; }
            run_phase<D, PROF, AG>(ph, src, src_pitch, jr, plan, red, warp, lane, fine, [&](int col, int seq, float2 lo, float2 hi) {
                if (kind == 0) {                   // q, k, v of this CTA's heads (bf16, local)
                    uint8_t* q = qkvs + (seq * 3 * HS + col) * 2;
                    *reinterpret_cast<uint32_t*>(q) = pack_bf16(lo.x, lo.y);
                    *reinterpret_cast<uint32_t*>(q + 16) = pack_bf16(hi.x, hi.y);
                } else if (kind == 4) {            // logits -> CTA 0 (fp32)
                    const uint32_t z = zbase + (seq * sm.zp + crank * VS + col) * 4;
                    if (seq < G) {
                        st_cluster_v2(z, __float_as_uint(lo.x), __float_as_uint(lo.y));
                        st_cluster_v2(z + 32, __float_as_uint(hi.x), __float_as_uint(hi.y));
                    }
                } else {
                    const int off = seq * dst_pitch + (col0 + col) * 2;
                    if (kind == 2) {               // gelu(c_fc)
                        lo.x = gelu_tanh(lo.x); lo.y = gelu_tanh(lo.y); hi.x = gelu_tanh(hi.x); hi.y = gelu_tanh(hi.y);
                    } else {                       // x2 = x1 + c_proj(att)  |  out = x2 + c_proj(gelu)
                        const float2 r0 = unpack_bf16(*reinterpret_cast<const uint32_t*>(res + off));
                        const float2 r1 = unpack_bf16(*reinterpret_cast<const uint32_t*>(res + off + 16));
                        lo.x += r0.x; lo.y += r0.y; hi.x += r1.x; hi.y += r1.y;
                    }
                    if (AG) {
                        // the tile (8 sequences x 16 columns) goes into this CTA's copy first; then it leaves as 16-byte
                        // pieces (lane = sequence and half of the 32-byte row): a quarter of the packets and of the
                        // byte-count updates of the mbarriers that single words would need
                        if (seq < G) {
                            *reinterpret_cast<uint32_t*>(dst + off) = pack_bf16(lo.x, lo.y);
                            *reinterpret_cast<uint32_t*>(dst + off + 16) = pack_bf16(hi.x, hi.y);
                        }
                        __syncwarp();
                        if (lane < 16 && (lane >> 1) < G)
                            broadcast16_async<CL>(dst + (lane >> 1) * dst_pitch + (col0 + col - (g & ~1)) * 2 + (lane & 1) * 16, gbar_a + 8 * kind);
                    } else {
                        broadcast_u32<CL>(dst + off, pack_bf16(lo.x, lo.y), seq < G);
                        broadcast_u32<CL>(dst + off + 16, pack_bf16(hi.x, hi.y), seq < G);
                    }
                }
            });
            if (kind == 4) break;                  // the step ends with the sampling below
            if (kind == 0) {
                __syncthreads();
                MG_PROF(2)
                // ---- P3: append k, v; attention of (sequence, head) pairs on the tensor cores; all-gather into Y ----
                {
                    MG_WAIT_SLOT(1)
                    __nv_bfloat16* cache_l = a.cache + static_cast<size_t>(l) * a.layer_stride;
                    const int nmine = plan.nmine;
                    const int nchunks = (pos + CT - 1) / CT;
                    const int ss = plan.ss;
                    const bool lead = plan.split == 0;             // appends k|v, seeds the softmax, writes the output
                    // ldmatrix row addresses of this lane inside a 16-token tile of k|v records: the K rows are the A
                    // operand of the score MMA (16 tokens x d_h), the V rows the B operand of P.V (transposed load)
                    // (pieces are swizzled, see kv_piece_off; the swizzle of a lane's row does not depend on the tile)
                    const int mi = lane >> 3, mr = lane & 7;
                    constexpr int PR = D / 4;
                    uint32_t k_lane[D / 16], v_lane[D / 16];
    #pragma unroll
                    for (int ks = 0; ks < D / 16; ++ks) {
                        k_lane[ks] = kv_piece_off<D>((mi & 1) * 8 + mr, ks * 2 + (mi >> 1));
                        // (V rows in the order the probabilities arrive in: k = 2t, 2t + 1, 2t + 8, 2t + 9 of the P.V MMA are
                        // tokens 2t, 2t + 8, 2t + 1, 2t + 9 of the tile, so that a lane's packed pair {token g, token g + 8}
                        // is an A-operand register as it stands)
                        v_lane[ks] = kv_piece_off<D>((mr & ~1) + (mi & 1) + (mr & 1) * 8, PR / 2 + ks * 2 + (mi >> 1));
                    }
                    FineProf afine{nullptr, 0};
                    if (PROF && profiling) { afine.acc = prof_acc + 54; afine.t = clock64(); }
                    for (int pi = 0; pi < nmine; ++pi) {
                        const int q = (warp + MG_WARPS * pi) >> ss;
                        const int sl = q / HPC, hh = q % HPC;
                        const int b = s0 + sl, h = crank * HPC + hh;
                        const uint32_t* qw = reinterpret_cast<const uint32_t*>(qkvs + (sl * 3 * HS + hh * D) * 2);
                        const uint32_t* kw = qw + HS / 2;
                        const uint32_t* vw = qw + HS;
                        // Scores as K q: the 16 tokens of a tile are the MMA rows and every column of the B operand is
                        // the query, so lane (g, tig) gets the scores of tokens g and g + 8 (the same in its 4 tig
                        // lanes: the softmax below runs on 16 distinct values per tile, not on 8 copies of the row).
                        // New token = the softmax seed.
                        uint32_t qb[D / 16][2];
                        float snew = 0.f;
    #pragma unroll
                        for (int ks = 0; ks < D / 16; ++ks) {
                            const uint32_t q0w = qw[ks * 8 + tig], q1w = qw[ks * 8 + 4 + tig];
                            qb[ks][0] = q0w;
                            qb[ks][1] = q1w;
                            const float2 q0 = unpack_bf16(q0w), q1 = unpack_bf16(q1w);
                            const float2 k0 = unpack_bf16(kw[ks * 8 + tig]), k1 = unpack_bf16(kw[ks * 8 + 4 + tig]);
                            snew += q0.x * k0.x + q0.y * k0.y + q1.x * k1.x + q1.y * k1.y;
                        }
                        snew += __shfl_xor_sync(0xffffffffu, snew, 1);
                        snew += __shfl_xor_sync(0xffffffffu, snew, 2);
                        float m = lead ? snew * a.scale_log2 : -INFINITY;        // running maximum (log2 units), warp-uniform
                        float lsum = (lead && lane < 4) ? 1.f : 0.f;    // this lane's share of 4 x the denominator (4 tig copies)
                        float o[NT_O][4];                               // row 0 of P.V; every quad carries a copy
    #pragma unroll
                        for (int dt = 0; dt < NT_O; ++dt) {
                            const float2 vn = unpack_bf16(vw[dt * 4 + tig]);
                            o[dt][0] = lead ? vn.x : 0.f; o[dt][1] = lead ? vn.y : 0.f; o[dt][2] = 0.f; o[dt][3] = 0.f;
                        }
                        // append the k|v record to the global cache
                        if (lead && lane < 2 * CH) {
                            const int part = lane % CH;
                            // (the pieces of a record are stored swizzled, so that a chunk is a linear image of a stage)
                            const size_t rec = ((static_cast<size_t>(b) * H + h) * a.t_max + pos) * (2 * D);
                            const uint4 val = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(lane < CH ? kw : vw) + part * 16);
                            *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(cache_l + rec) + kv_piece_off<D>(pos % CT, lane) - (pos % CT) * REC) = val;
                            asm volatile("fence.proxy.async.global;" ::: "memory");    // read back by TMA in a later job
                        }
                        __syncwarp();
                        if (PROF) afine.mark(0);
                        for (int c = plan.split; c < nchunks; c += 1 << ss) {
                            uint8_t* st = ring_acquire<PROF>(jr);
                            const int ntok = min(CT, pos - c * CT);
                            const uint32_t st_a = smem_u32(st);
                            // S^T = K q: s[j][0] = token 16 j + g, s[j][2] = token 16 j + 8 + g (lanes with tig == 0)
                            float s[CT / 16][4];
    #pragma unroll
                            for (int jt = 0; jt < CT / 16; ++jt) {
                                s[jt][0] = s[jt][1] = s[jt][2] = s[jt][3] = 0.f;
    #pragma unroll
                                for (int ks = 0; ks < D / 16; ++ks) {
                                    uint32_t ka[4];
                                    ldmatrix_x4(ka, st_a + jt * 16 * REC + k_lane[ks]);
                                    mma_16816(s[jt], ka[0], ka[1], ka[2], ka[3], qb[ks][0], qb[ks][1]);
                                }
                            }
                            // EARLY (full clusters: the attention phase is paced by the ring): the V rows of the chunk move
                            // to registers now and the stage is re-armed before the softmax, ~1 k cycles earlier, so
                            // that more bytes are in flight per warp (232.7 vs 240.5 us per step at 256 sequences).  With few
                            // sequences per cluster the phase is a latency chain and the re-arm is better placed behind the
                            // P V MMAs, whose latency it overlaps (111.1 vs 107.8 us at 32 sequences, 129.5 vs 123.8 at 64).
                            // (The XOR makes the re-arm wait for the loads: an ldmatrix has finished reading the stage when
                            // its registers can be read.)
                            constexpr bool EARLY = !AG;
                            uint32_t vb[CT / 16][D / 16][4];
                            if (EARLY) {
                                uint32_t vdep = 0;
    #pragma unroll
                                for (int jt = 0; jt < CT / 16; ++jt)
    #pragma unroll
                                    for (int dp = 0; dp < D / 16; ++dp) {
                                        ldmatrix_x4_trans(vb[jt][dp], st_a + jt * 16 * REC + v_lane[dp]);
                                        vdep ^= vb[jt][dp][0] ^ vb[jt][dp][1] ^ vb[jt][dp][2] ^ vb[jt][dp][3];
                                    }
                                asm volatile("" ::"r"(vdep) : "memory");
                                ring_release<D>(jr, plan, lane);
                            }
                            if (ntok < CT) {       // last, partial chunk of the pair: tokens that are not cached yet
    #pragma unroll
                                for (int jt = 0; jt < CT / 16; ++jt) {
                                    if (jt * 16 + g >= ntok) s[jt][0] = -INFINITY;
                                    if (jt * 16 + 8 + g >= ntok) s[jt][2] = -INFINITY;
                                }
                            }
                            float tm[CT / 16];                 // (a tree: the chain of dependent maxima was on the critical path)
    #pragma unroll
                            for (int jt = 0; jt < CT / 16; ++jt) tm[jt] = fmaxf(s[jt][0], s[jt][2]);
    #pragma unroll
                            for (int w = CT / 32; w > 0; w >>= 1)
    #pragma unroll
                                for (int jt = 0; jt < w; ++jt) tm[jt] = fmaxf(tm[jt], tm[jt + w]);
                            float mx = tm[0];
                            mx = warp_max_f32(mx);              // one CREDUX instead of three shuffle + max rounds
                            const float mn = fmaxf(m, mx * a.scale_log2);
                            const float corr = fast_exp2(m - mn);
                            m = mn;
                            lsum *= corr;
    #pragma unroll
                            for (int dt = 0; dt < NT_O; ++dt) { o[dt][0] *= corr; o[dt][1] *= corr; }
                            // O += P V: the probabilities of 16 tokens travel to the A-operand layout (row 0 = p over k)
                            // as bf16 pairs: lane (g, tig) needs tokens 2 tig, 2 tig + 1 (from lanes 8 tig, 8 tig + 4)
    #pragma unroll
                            for (int jt = 0; jt < CT / 16; ++jt) {
                                const float p_lo = fast_exp2(fmaf(s[jt][0], a.scale_log2, -mn));
                                const float p_hi = fast_exp2(fmaf(s[jt][2], a.scale_log2, -mn));
                                lsum += p_lo + p_hi;
                                const uint32_t pk = pack_bf16(p_lo, p_hi);            // {token g, token g + 8} of this tile
                                const uint32_t a0 = __shfl_sync(0xffffffffu, pk, 8 * tig);        // tokens 2 tig, 2 tig + 8
                                const uint32_t a2 = __shfl_sync(0xffffffffu, pk, 8 * tig + 4);    // tokens 2 tig + 1, 2 tig + 9
    #pragma unroll
                                for (int dp = 0; dp < D / 16; ++dp) {
                                    if (!EARLY) ldmatrix_x4_trans(vb[jt][dp], st_a + jt * 16 * REC + v_lane[dp]);
                                    mma_16816(o[2 * dp], a0, 0u, a2, 0u, vb[jt][dp][0], vb[jt][dp][1]);
                                    mma_16816(o[2 * dp + 1], a0, 0u, a2, 0u, vb[jt][dp][2], vb[jt][dp][3]);
                                }
                            }
                            if (!EARLY) ring_release<D>(jr, plan, lane);
                        }
                        if (PROF) afine.mark(1);
                        lsum += __shfl_xor_sync(0xffffffffu, lsum, 4);
                        lsum += __shfl_xor_sync(0xffffffffu, lsum, 8);
                        lsum += __shfl_xor_sync(0xffffffffu, lsum, 16);
                        lsum += __shfl_xor_sync(0xffffffffu, lsum, 1);
                        lsum += __shfl_xor_sync(0xffffffffu, lsum, 2);
                        if (ss > 0) {                                   // merge the partial states of the pair's warps
                            constexpr int PW = 68;                      // floats per warp: m, l, -, -, o[D]
                            float* part = red + warp * PW;
                            const uint32_t bar_id = 1 + (warp >> ss), bar_n = 32u << ss;
                            if (!lead) {
                                if (lane == 0) { part[0] = m; part[1] = lsum; }
                                if (lane < 4) {
    #pragma unroll
                                    for (int dt = 0; dt < NT_O; ++dt) *reinterpret_cast<float2*>(part + 4 + dt * 8 + 2 * tig) = make_float2(o[dt][0], o[dt][1]);
                                }
                            }
                            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_n) : "memory");
                            if (lead) {
                                float mm = m;
                                for (int j = 1; j < (1 << ss); ++j) mm = fmaxf(mm, part[j * PW]);
                                const float w0 = fast_exp2(m - mm);
                                lsum *= w0;
    #pragma unroll
                                for (int dt = 0; dt < NT_O; ++dt) { o[dt][0] *= w0; o[dt][1] *= w0; }
                                for (int j = 1; j < (1 << ss); ++j) {
                                    const float* pj = part + j * PW;
                                    const float wj = fast_exp2(pj[0] - mm);
                                    lsum = fmaf(wj, pj[1], lsum);
    #pragma unroll
                                    for (int dt = 0; dt < NT_O; ++dt) {
                                        const float2 oj = *reinterpret_cast<const float2*>(pj + 4 + dt * 8 + 2 * tig);
                                        o[dt][0] = fmaf(wj, oj.x, o[dt][0]);
                                        o[dt][1] = fmaf(wj, oj.y, o[dt][1]);
                                    }
                                }
                            }
                            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_n) : "memory");      // the slots are free again
                        }
                        if (PROF) afine.mark(2);
                        const float inv = 4.0f / lsum;                  // lsum counted every token in its 4 tig lanes
                        // every quad holds the same output row: quad g sends it to CTA g of the cluster
                        if (AG) {
                            if (lead) {
                                // lane tig of a quad collects the 16-byte pieces tig, tig + 4 of the head's output row (every
                                // quad holds the whole row); quad g sends them to CTA g
                                constexpr int NP = (NT_O + 3) / 4;
                                uint4 piece[NP] = {};
    #pragma unroll
                                for (int dt = 0; dt < NT_O; ++dt) {
                                    const uint32_t pk = pack_bf16(o[dt][0] * inv, o[dt][1] * inv);
                                    const uint32_t w0 = __shfl_sync(0xffffffffu, pk, lane & ~3), w1 = __shfl_sync(0xffffffffu, pk, (lane & ~3) + 1);
                                    const uint32_t w2 = __shfl_sync(0xffffffffu, pk, (lane & ~3) + 2), w3 = __shfl_sync(0xffffffffu, pk, (lane & ~3) + 3);
                                    if (tig == (dt & 3)) piece[dt >> 2] = make_uint4(w0, w1, w2, w3);
                                }
                                if (g < CL) {
                                    const uint32_t rb = map_to_cta(gbar_a, g);
                                    const uint32_t dst = rb + (smem_u32(Y + sl * pe + h * D * 2) - gbar_a);
    #pragma unroll
                                    for (int i = 0; i < NP; ++i)
                                        if (4 * i + tig < NT_O) st_async_v4(dst + (4 * i + tig) * 16, piece[i], rb);
                                }
                            }
                        } else if (lead && g < CL) {
                            const uint32_t dst = map_to_cta(smem_u32(Y + sl * pe + (h * D + 2 * tig) * 2), g);
    #pragma unroll
                            for (int dt = 0; dt < NT_O; ++dt) st_cluster_u32(dst + dt * 16, pack_bf16(o[dt][0] * inv, o[dt][1] * inv));
                        }
                        if (PROF) afine.mark(3);
                    }
                }
                if (use_ln) ln_prefetch(lnf, P + lw.ln2_g, P + lw.ln2_b, E, lane);
            }
            MG_PROF(kind == 0 ? 3 : 2 * kind + 3)
            if (AG) {                              // every slice of the all-gathered activation has arrived in this CTA
                // (one warp polls the mbarrier, the others sleep at the CTA barrier)
                if (warp == 0) {
                    mbar_wait(&gbar[kind], gpar);
                    if (lane == 0) mbar_expect_tx(&gbar[kind], gbytes[kind]);
                }
                __syncthreads();
                if (kind == 3) gpar ^= 1u;
            } else {
                cluster_sync_all();                // the all-gathered activation is complete in every CTA
            }
            MG_PROF(2 * kind + 4)
            if (kind == 3) { uint8_t* t = X; X = Y; Y = t; }        // the block output becomes the next block's input
        }
        MG_PROF(11)
        cluster_sync_all();                        // E: all logits are in CTA 0
        MG_PROF(12)
        if (crank == 0 && warp < G) {
            const int b = s0 + warp;
            float u;
            float* zrow = Z + warp * sm.zp;
            if (a.logits_out != nullptr && step == a.steps - 1)
                for (int c = lane; c < V; c += 32) a.logits_out[static_cast<size_t>(b) * V + c] = zrow[c];
            __syncwarp();
            int chosen = sample_row_smem(zrow, V, a.inv_temperature, a.greedy, a.seed_lo, a.seed_hi,
                                         static_cast<uint32_t>(a.seq_base + b), static_cast<uint32_t>(step), lane, &u);
            if (a.forced != nullptr) {
                const int f = a.forced[static_cast<size_t>(b) * a.steps + step];
                if (f >= 0) chosen = f;
            }
            if (lane == 0) {
                a.out_ids[static_cast<size_t>(b) * a.steps + step] = chosen;
                if (a.uniforms != nullptr) a.uniforms[static_cast<size_t>(b) * a.steps + step] = u;
            }
            if (lane < CL) st_cluster_u32(map_to_cta(smem_u32(&toks[warp]), lane), static_cast<uint32_t>(chosen));
        }
        MG_PROF(13)
        cluster_sync_all();                        // F: next tokens are everywhere
        MG_PROF(14)
    }
    if (profiling)
        for (int i = 0; i < 64; ++i) a.prof[i] = prof_acc[i];
#undef MG_PROF
#undef MG_WAIT_SLOT
}

// ---------------------------------------------------------------------------
// Weight stream: one entry per linear layer, in the order the kernel consumes them.
// ---------------------------------------------------------------------------
struct PackPhase {
    long long w_src;      // bf16 [N, K] matrix in the shadow arena
    long long b_src;      // fp32 bias in the parameter arena, or -1
    int K, ntiles, ksplit, tpp, pstride, cols_per_cta, row_limit, sub, slot0, nslots;
};

// One block per (slot, CTA rank): 256 16-byte weight chunks in A-fragment order + 4 chunks of bias.  Chunk
// (i, m, lane = (g, tig)) of a slot holds what lane needs for MMA m of k-block i: weight rows r0 + g and r0 + g + 8
// at k = kbase + 32 i + 8 tig + 4 m + {0, 1} and + {2, 3}.
__global__ void __launch_bounds__(288)
mega_pack_kernel(const __nv_bfloat16* __restrict__ shadow, const float* __restrict__ params, uint8_t* __restrict__ stream,
                 const PackPhase* __restrict__ table, int nphase, int per_step) {
    const int s = blockIdx.x, c = blockIdx.y, t = threadIdx.x;
    if (t >= 260) return;
    int pi = 0;
    while (pi + 1 < nphase && s >= table[pi + 1].slot0) ++pi;
    const PackPhase ph = table[pi];
    const int local = s - ph.slot0;
    const int u = local / ph.sub, j = local % ph.sub;
    const int nt = u % ph.ntiles, ks = u / ph.ntiles;
    const int row0 = c * ph.cols_per_cta + (nt / ph.tpp) * ph.pstride + (nt % ph.tpp) * 16;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (t < 256) {
        const int i = t >> 6, m = (t >> 5) & 1, lane = t & 31, g = lane >> 2, tig = lane & 3;
        const int k = (ks * ph.sub + j) * 128 + 32 * i + 8 * tig + 4 * m;
        const uint32_t* lo = reinterpret_cast<const uint32_t*>(shadow + ph.w_src + static_cast<long long>(row0 + g) * ph.K + k);
        const uint32_t* hi = reinterpret_cast<const uint32_t*>(shadow + ph.w_src + static_cast<long long>(row0 + g + 8) * ph.K + k);
        if (row0 + g <= ph.row_limit) { val.x = lo[0]; val.z = lo[1]; }
        if (row0 + g + 8 <= ph.row_limit) { val.y = hi[0]; val.w = hi[1]; }
    } else if (ph.b_src >= 0 && ks == 0 && j == 0) {
        float b[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int row = row0 + (t - 256) * 4 + e;
            b[e] = row <= ph.row_limit ? params[ph.b_src + row] : 0.f;
        }
        val = make_uint4(__float_as_uint(b[0]), __float_as_uint(b[1]), __float_as_uint(b[2]), __float_as_uint(b[3]));
    }
    *reinterpret_cast<uint4*>(stream + (static_cast<size_t>(c) * per_step + s) * MG_SLOT + t * 16) = val;
}

// K split of a phase: fill the 16 warps (one unit each) with pieces of at least 128 of K.
static int mega_ksplit(int ntiles, int K) {
    int ks = 1;
    while (ntiles * ks * 2 <= MG_WARPS && K / (ks * 2) >= 128) ks *= 2;
    return ks;
}

static MegaSmem mega_smem_layout(int E, int F, int V, int D, int CL, int nst, int nst_extra) {
    MegaSmem s{};
    const int HS = E / CL, FS = F / CL;
    s.vs = 16 * ((V + 16 * CL - 1) / (16 * CL));
    s.pe = 2 * E + 64;
    s.pf = 2 * F + 64;
    s.ks_attn = mega_ksplit(3 * HS / 16, E);
    s.ks_proj = mega_ksplit(HS / 16, E);
    s.ks_fc = mega_ksplit(FS / 16, E);
    s.ks_proj2 = mega_ksplit(HS / 16, F);
    s.ks_logits = mega_ksplit(s.vs / 16, E);
    s.n_attn = (3 * HS / 16) * (E / 128);
    s.n_proj = (HS / 16) * (E / 128);
    s.n_fc = (FS / 16) * (E / 128);
    s.n_proj2 = (HS / 16) * (F / 128);
    s.n_logits = (s.vs / 16) * (E / 128);
    s.per_layer = s.n_attn + s.n_proj + s.n_fc + s.n_proj2;
    int off = 0;
    auto take = [&](int bytes) { int o = off; off = (off + bytes + 127) & ~127; return o; };
    s.buf0 = take(MG_ROWS * s.pe);
    s.buf1 = take(MG_ROWS * s.pe);
    s.bufn = take(MG_ROWS * s.pe);
    s.bufg = take(MG_ROWS * s.pf);
    s.qkv = take(MG_ROWS * 3 * HS * 2);
    s.red = take(MG_WARPS * 32 * 16);      // K-split partials: <= 15 units x 32 lanes x 4 floats
    s.jobtab = take(MG_WARPS * MG_TAB * 2);
    s.nst = nst;
    s.nst_extra = nst_extra;
    s.zp = CL * s.vs;
    // The logits matrix (CTA 0) shares the GELU buffer: c_fc writes that buffer and mlp c_proj reads it before the
    // barrier in front of the logits phase, and the sampler has read the logits before the barrier that ends the
    // step; every kilobyte not spent here is ring depth, which the attention phase is sensitive to.
    if (MG_ROWS * s.zp * 4 <= MG_ROWS * s.pf) s.zbuf = s.bufg;
    else s.zbuf = take(MG_ROWS * s.zp * 4);
    s.ring = take((MG_WARPS * nst + nst_extra) * MG_STAGE);
    s.bars = take((MG_WARPS * nst + nst_extra + 4) * 8);      // ring stages, then the four all-gather barriers
    s.toks = take(MG_ROWS * 4);
    s.prof = take(64 * 8);
    s.total = off;
    return s;
}

// Ring stages: as many as fit, at most 4 per warp (every stage is a 4 KB job in flight); when the number that fits
// is not a multiple of 16 the first warps get one more.  A job is issued `stages` jobs ahead of its use and the KV
// jobs of a step must not be issued before the previous step appended to the cache, which the >= 2 weight jobs per
// layer and warp guarantee only from 2 layers on: a 1-layer model is limited to 2 stages.
static MegaSmem mega_smem_fit(int E, int V, int D, int CL, int L) {
    constexpr int LIMIT = 227 * 1024;
    int max_stages = (L >= 2 ? 4 : 2) * MG_WARPS;
    if (const char* env = getenv("CB200_DECODE_RING_STAGES")) {      // tuning knob: stages per warp
        const int v = atoi(env);
        if (v >= 2 && v * MG_WARPS < max_stages) max_stages = v * MG_WARPS;
    }
    const MegaSmem base = mega_smem_layout(E, 4 * E, V, D, CL, 0, 0);
    int stages = (LIMIT - base.total - 512) / (MG_STAGE + 8);
    if (stages > max_stages) stages = max_stages;
    if (stages < 2 * MG_WARPS) stages = 2 * MG_WARPS;                 // does not fit: reported through .total
    MegaSmem sm = mega_smem_layout(E, 4 * E, V, D, CL, stages / MG_WARPS, stages % MG_WARPS);
    sm.per_step = L * sm.per_layer + sm.n_logits;
    return sm;
}

template <int D, int CL, bool PROF = false, bool AG = true>
static int mega_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, const MegaSmem& sm) {
    auto kernel = decode_mega_kernel<D, CL, PROF, AG>;
    static int configured_smem = 0;
    if (configured_smem < sm.total) {
        CB200_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sm.total));
        configured_smem = sm.total;
    }
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(MG_THREADS);
    cfg.gridDim = dim3(CL);
    cfg.dynamicSmemBytes = sm.total;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return 0;
}

template <int D, int CL>
static int mega_capacity(const MegaSmem& sm) {
    static int cached = -1, cached_smem = -1;
    if (cached >= 0 && cached_smem == sm.total) return cached;
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    if (mega_config<D, CL>(cfg, attr, sm)) return -1;
    int resident = 0;
    if (cudaOccupancyMaxActiveClusters(&resident, decode_mega_kernel<D, CL>, &cfg) != cudaSuccess) return -1;
    cached = resident; cached_smem = sm.total;
    return resident;
}

// Clusters to launch: one wave when the batch fits (<= 8 sequences per cluster), otherwise 8-sequence clusters in
// several waves.
static int mega_cluster_count(int B, int resident, int max_clusters_hint) {
    if (max_clusters_hint > 0 && max_clusters_hint < resident) resident = max_clusters_hint;
    int ncl = B < resident ? B : resident;
    if (static_cast<long long>(ncl) * MG_ROWS < B) ncl = (B + MG_ROWS - 1) / MG_ROWS;
    // the fewest clusters with the same largest share of sequences: a cluster's step time hardly depends on its
    // share, the slowest cluster sets the time, and clusters that are not needed only add L2 traffic and the risk
    // that the last one does not become resident with the others (it would then run after one of them has finished)
    const int share = (B + ncl - 1) / ncl;
    return (B + share - 1) / share;
}

template <int D, int CL, bool PROF = false, bool AG = true>
static int launch_mega(const MegaArgs& args, const MegaSmem& sm, int ncl, cudaStream_t s) {
    if (!PROF && args.prof != nullptr) return launch_mega<D, CL, true, AG>(args, sm, ncl, s);    // diagnostic instantiation
    if (AG && !args.async_gather) return launch_mega<D, CL, PROF, false>(args, sm, ncl, s);      // cluster barriers
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    int rc = mega_config<D, CL, PROF, AG>(cfg, attr, sm);
    if (rc) return rc;
    cfg.stream = s;
    cfg.gridDim = dim3(ncl * CL);
    CB200_CUDA_OK(cudaLaunchKernelEx(&cfg, decode_mega_kernel<D, CL, PROF, AG>, args, sm));
    note_launch(1);
    return 0;
}

static bool mega_shape_ok(int E, int H, int D, int V, int L, int CL) {
    if (!(E == 256 || E == 512)) return false;
    if (H % CL != 0 || H * D != E) return false;
    if (!(D == 16 || D == 32 || D == 64)) return false;
    if (L < 1 || L > MG_MAX_LAYERS || V < 1 || V > 4096) return false;
    if ((E / CL) % 16 != 0) return false;
    const MegaSmem sm = mega_smem_fit(E, V, D, CL, L);
    return sm.total <= 227 * 1024;
}

bool decode_mega_supported(int E, int H, int D, int V, int L) {
    return mega_shape_ok(E, H, D, V, L, 8) || mega_shape_ok(E, H, D, V, L, 4);
}

// Bytes of the packed weight stream (+ its phase table) for the larger of the two cluster sizes.
int64_t decode_mega_stream_bytes(int E, int H, int D, int V, int L) {
    int64_t need = 0;
    for (int CL = 4; CL <= 8; CL += 4) {
        if (!mega_shape_ok(E, H, D, V, L, CL)) continue;
        const MegaSmem sm = mega_smem_fit(E, V, D, CL, L);
        const int64_t bytes = static_cast<int64_t>(CL) * sm.per_step * MG_SLOT + (4 * MG_MAX_LAYERS + 1) * sizeof(PackPhase) + 256;
        if (bytes > need) need = bytes;
    }
    return need;
}

#define CB200_MEGA_DISPATCH(FN, ...)                                            \
    (CL == 8 ? (D == 16 ? FN<16, 8>(__VA_ARGS__) : D == 32 ? FN<32, 8>(__VA_ARGS__) : FN<64, 8>(__VA_ARGS__)) \
             : (D == 16 ? FN<16, 4>(__VA_ARGS__) : D == 32 ? FN<32, 4>(__VA_ARGS__) : FN<64, 4>(__VA_ARGS__)))

// Clusters of `CL` CTAs of the persistent decode kernel that are co-resident on the current device.
int decode_mega_capacity(int E, int H, int V, int D, int L, int CL) {
    if (!mega_shape_ok(E, H, D, V, L, CL)) return 0;
    const MegaSmem sm = mega_smem_fit(E, V, D, CL, L);
    return CB200_MEGA_DISPATCH(mega_capacity, sm);
}

// ---- batched prefill: the k|v records of T prompt tokens in this kernel's cache layout ----
// One thread per 16-byte piece of a record (pieces 0 .. D/8-1 = k, the rest = v), written where the attention phase's
// append would have put it (kv_piece_off: the swizzle inside the CT-token chunk).
template <int D>
__global__ void __launch_bounds__(256)
kv_export_mega_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ cache_l, int B, int T, int H, int t_max) {
    constexpr int CT = MG_CT(D), REC = 4 * D, CH = D / 8;
    const int E = H * D;
    const size_t n = static_cast<size_t>(B) * T * H * 2 * CH;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        size_t r = i;
        const int piece = static_cast<int>(r % (2 * CH)); r /= 2 * CH;
        const int h = static_cast<int>(r % H); r /= H;
        const int t = static_cast<int>(r % T);
        const int b = static_cast<int>(r / T);
        const int part = piece % CH, kv = piece / CH;
        const uint4 val = *reinterpret_cast<const uint4*>(qkv + (static_cast<size_t>(b) * T + t) * 3 * E + (1 + kv) * E + h * D + part * 8);
        const size_t rec = ((static_cast<size_t>(b) * H + h) * t_max + t) * (2 * D);
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(cache_l + rec) + kv_piece_off<D>(t % CT, piece) - (t % CT) * REC) = val;
    }
}

int kv_export_mega(const __nv_bfloat16* qkv, __nv_bfloat16* cache_layer, int B, int T, int H, int D, int t_max,
                   cudaStream_t s) {
    const size_t n = static_cast<size_t>(B) * T * H * (D / 4);
    const int blocks = static_cast<int>(n / 256 + 1 < 148 * 8 ? n / 256 + 1 : 148 * 8);
    switch (D) {
        case 16: kv_export_mega_kernel<16><<<blocks, 256, 0, s>>>(qkv, cache_layer, B, T, H, t_max); break;
        case 32: kv_export_mega_kernel<32><<<blocks, 256, 0, s>>>(qkv, cache_layer, B, T, H, t_max); break;
        case 64: kv_export_mega_kernel<64><<<blocks, 256, 0, s>>>(qkv, cache_layer, B, T, H, t_max); break;
        default: set_error("head size %d is not supported", D); return -1;
    }
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

// cluster_size 0 = automatic: 8-CTA clusters (each weight is read by fewer CTAs) when the batch fits one wave of
// them, otherwise 4-CTA clusters (more of them are co-resident: GPCs rarely have a multiple of 8 SMs free).
// `stream_ws` receives the packed weight stream (decode_mega_stream_bytes).
int decode_mega(MegaArgs args, int D, int max_clusters, int cluster_size, uint8_t* stream_ws, int64_t stream_ws_bytes,
                cudaStream_t s) {
    CB200_REQUIRE(decode_mega_supported(args.E, args.H, D, args.V, args.L), "shape not supported by the cluster decode kernel");
    CB200_REQUIRE(cluster_size == 0 || cluster_size == 4 || cluster_size == 8, "cluster size must be 0 (auto), 4 or 8");
    if (args.B == 0 || args.steps == 0) return 0;
    int CL = cluster_size;
    if (CL == 0) {
        const int cap8 = decode_mega_capacity(args.E, args.H, args.V, D, args.L, 8);
        const int cap4 = decode_mega_capacity(args.E, args.H, args.V, D, args.L, 4);
        CL = (cap8 > 0 && (args.B <= cap8 * MG_ROWS || cap4 <= 0)) ? 8 : 4;
    }
    CB200_REQUIRE(mega_shape_ok(args.E, args.H, D, args.V, args.L, CL), "cluster size %d does not fit this shape", CL);
    const int resident = decode_mega_capacity(args.E, args.H, args.V, D, args.L, CL);
    CB200_REQUIRE(resident >= 1, "the cluster decode kernel does not fit on this device");
    if (const char* env = getenv("CB200_DECODE_MAX_CLUSTERS")) {       // tuning knob
        const int v = atoi(env);
        if (v > 0 && (max_clusters == 0 || v < max_clusters)) max_clusters = v;
    }
    const int ncl = mega_cluster_count(args.B, resident, max_clusters);
    const MegaSmem sm = mega_smem_fit(args.E, args.V, D, CL, args.L);
    // ---- pack the weight stream ----
    const int E = args.E, F = args.F, HS = E / CL, FS = F / CL;
    std::vector<PackPhase> table;
    int slot = 0;
    auto add = [&](long long w, long long b, int K, int ntiles, int ksplit, int tpp, int pstride, int cols, int limit, int sub) {
        PackPhase p{w, b, K, ntiles, ksplit, tpp, pstride, cols, limit, sub, slot, ntiles * ksplit * sub};
        slot += p.nslots;
        table.push_back(p);
    };
    for (int l = 0; l < args.L; ++l) {
        const MegaLayer& w = args.layers[l];
        add(w.attn_w, w.attn_b, E, 3 * HS / 16, sm.ks_attn, HS / 16, E, HS, 3 * E - 1, E / 128 / sm.ks_attn);
        add(w.proj_w, w.proj_b, E, HS / 16, sm.ks_proj, HS / 16, 0, HS, E - 1, E / 128 / sm.ks_proj);
        add(w.fc_w, w.fc_b, E, FS / 16, sm.ks_fc, FS / 16, 0, FS, F - 1, E / 128 / sm.ks_fc);
        add(w.proj2_w, w.proj2_b, F, HS / 16, sm.ks_proj2, HS / 16, 0, HS, E - 1, F / 128 / sm.ks_proj2);
    }
    add(args.wte_sh, -1, E, sm.vs / 16, sm.ks_logits, sm.vs / 16, 0, sm.vs, args.V - 1, E / 128 / sm.ks_logits);
    CB200_REQUIRE(slot == sm.per_step, "weight stream plan mismatch: %d slots, expected %d", slot, sm.per_step);
    const int64_t stream_bytes = static_cast<int64_t>(CL) * sm.per_step * MG_SLOT;
    const int64_t table_off = (stream_bytes + 255) & ~int64_t(255);
    CB200_REQUIRE(stream_ws != nullptr && stream_ws_bytes >= table_off + static_cast<int64_t>(table.size() * sizeof(PackPhase)),
                  "weight stream workspace too small");
    CB200_CUDA_OK(cudaMemcpyAsync(stream_ws + table_off, table.data(), table.size() * sizeof(PackPhase), cudaMemcpyHostToDevice, s));
    CB200_CUDA_OK(cudaStreamSynchronize(s));       // `table` is pageable host memory that dies with this frame
    mega_pack_kernel<<<dim3(sm.per_step, CL), 288, 0, s>>>(args.shadow, args.params, stream_ws,
                                                            reinterpret_cast<const PackPhase*>(stream_ws + table_off),
                                                            static_cast<int>(table.size()), sm.per_step);
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    args.wstream = stream_ws;
    args.kv_split_log2 = 2;
    if (const char* env = getenv("CB200_DECODE_KV_SPLIT")) args.kv_split_log2 = std::min(3, std::max(0, atoi(env)));      // tuning knob
    args.l2_hints = 1;
    if (const char* env = getenv("CB200_DECODE_L2_HINTS")) args.l2_hints = atoi(env) != 0;
    // All-gathers through st.async + mbarriers unless the clusters are full: measured on one B200 (us per step,
    // st.async vs cluster barriers) 101.7 / 108.6 at 16 sequences, 107.8 / 114.6 at 32, 123.4 / 130.1 at 64,
    // 134.1 / 139.8 at 96 (8-CTA clusters), 156.6 / 157.4 at 128, 202.5 / 203.5 at 192, 245.1 / 242.6 at 256 (4-CTA
    // clusters of 4, 6, 8 sequences): each 16-byte st.async costs its destination ~2 cycles, which at 8 sequences per
    // cluster outweighs the four cluster barriers per block.
    args.async_gather = CL == 8 || (args.B + ncl - 1) / ncl <= 6;
    if (const char* env = getenv("CB200_DECODE_ASYNC_GATHER")) args.async_gather = atoi(env) != 0;      // tuning knob
    // L2 prefetch budget of an attention phase, dealt evenly to the warps of all CTAs as whole 4 KB chunks
    int prefetch_mb = 0;          // (measured: a net loss, the prefetch traffic slows the GEMM phases by more than its hits save)
    if (const char* env = getenv("CB200_DECODE_KV_PREFETCH_MB")) prefetch_mb = std::max(0, atoi(env));                  // tuning knob
    args.kv_prefetch = static_cast<int>(std::min<long long>(30, (static_cast<long long>(prefetch_mb) << 20) / (static_cast<long long>(ncl) * CL * MG_WARPS * MG_STAGE)));
    if (args.prof != nullptr) CB200_CUDA_OK(cudaMemsetAsync(args.prof, 0, 64 * sizeof(long long), s));
    return CB200_MEGA_DISPATCH(launch_mega, args, sm, ncl, s);
}

}  // namespace cb200
