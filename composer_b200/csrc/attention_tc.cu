// Attention backward on the 5th-generation tensor cores (tcgen05 / TMEM).
//
// One CTA owns 128 keys of one (batch, head) and walks the 128-row query tiles at or below the diagonal.  A
// tile is processed in two 64-key halves, so that S and dP of a half take 128 TMEM columns and two CTAs fit one SM
// (d_h 16: 192 of 256 columns each): their phases interleave, which is what keeps the MUFU and the issue slots
// busy -- a single CTA has all its softmax warps wait, exponentiate and store in lock step.  Per half:
//   S  = Q K^T, dP = dO V^T          tcgen05.mma, M 128 x N 64, accumulators in TMEM
//   softmax warps (thread = query row, 32 columns each) pull their S and dP columns into registers at once and
//     release the TMEM columns (the other half's MMAs are issued under this half's arithmetic), then
//     P = exp2(S c - lse), causal mask, dropout, dS' = P (M.dP - delta/ks) -> bf16 -> swizzled smem
// and per tile, once both halves are in shared memory:
//   dV += P^T dO, dK += dS'^T Q      the same smem tile read through an MN-major
//   dQ  = dS' K                       descriptor (transposed) and a K-major one
// so no score element is ever transposed or re-loaded by a CUDA core.  dK, dV accumulate in TMEM over the whole
// walk; dQ tiles are drained from TMEM by the thread that owns the row and reduced into the fp32 dq buffer.
// Element arithmetic: packed fp32 pairs (FFMA2 / FMUL2), one MUFU and ~6 instructions per score element.
//
// Q, K, V, dO tiles are rows of d_h bf16 (32 / 64 / 128 bytes) loaded by TMA with the swizzle whose span is one
// row, consumed K-major (S, dP) and MN-major (dV, dK, dQ right operands).  Reference semantics: attention_fwd_tc.cu.
#include <cstdlib>

#include "attention.h"
#include "gemm.h"

#include <type_traits>

namespace cb200 {

static long long* g_attention_trace = nullptr;   // device buffer of 12 x 512 words, see the kernel
void attention_set_trace(long long* buffer) { g_attention_trace = buffer; }
long long* attention_get_trace() { return g_attention_trace; }

// SW softmax warps per CTA: 8 = 4 row bands x 2 column slices of a 64-key half (32 keys per thread and hand-off),
// 4 = one warp per row band (64 keys per thread and hand-off: half the hand-offs and per-tile bookkeeping per element,
// half the warps).  Then one warpgroup of issuing warps: loads + S/dP, dV (+ TMEM alloc), dK, dQ.
constexpr int TCB_TILE = 128;                          // query rows per tile = keys per CTA

template <int D>
struct TcbCfg {
    static constexpr int RB = 2 * D;                       // bytes per row of a Q/K/V/dO tile
    static constexpr int TILE = TCB_TILE * RB;             // 4 / 8 / 16 KB
    static constexpr int PBYTES = 2 * TCB_TILE * 128;      // P or dS': [2 key halves][128 rows][128 B]
    static constexpr int CPS = (D == 16) ? 2 : 1;          // CTAs per SM
    static constexpr int NBUF_Q = (D == 64) ? 2 : 3;       // Q / dO tile ring
    static constexpr int NBUF_P = (D == 32) ? 2 : 1;       // P / dS' buffers
    static constexpr size_t SMEM = 2 * TILE + 2 * NBUF_Q * TILE + 2 * NBUF_P * PBYTES + 256 + 1024;
    static constexpr int TMEM_COLS = (128 + 4 * D <= 256) ? 256 : 512;
    // registers (CPS 2): launched with 80 per thread (768 threads per SM); the control warpgroup keeps 32 and what
    // it releases, 128 * 48, lets the 8 softmax warps grow to 104 (setmaxnreg.inc only takes what the CTA's own
    // warps released).  CPS 1: 168 per thread from the start, no redistribution.
    static constexpr int SOFTMAX_REGS = 104, CONTROL_REGS = 32;     // SW 8
    static constexpr int SOFTMAX_REGS4 = 224;                       // SW 4: 256 threads x 128 at launch, 128 * 96 released
    static_assert(CPS * TMEM_COLS <= 512, "TMEM budget");
};

// Software pipeline (tile it, half h; half-step n = 2 it + h):
//   control thread:  S/dP of the next half-step is issued as soon as the softmax warps have pulled the current
//                    one out of TMEM (bar_sdp_free); dV/dK/dQ(it) follow when both halves of P/dS'(it) are in
//                    shared memory (bar_p_full).
//   softmax warps:   wait S/dP(n) -> tcgen05.ld -> release TMEM -> arithmetic -> smem -> bar_p_full;
//                    dQ(it-1) is drained at the end of tile it, when its MMAs have long finished.
// TRACED: the event timeline is compiled in (a separate instantiation, see attention_fwd_tc.cu).
template <int D, bool DROP, int SW = 8, bool TRACED = false>
__global__ void __launch_bounds__((SW + 4) * 32, TcbCfg<D>::CPS)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                   const float* __restrict__ lse, const float* __restrict__ delta, float* __restrict__ dq_acc,
                   __nv_bfloat16* __restrict__ dqkv, int T, int H, float scale, float scale_log2, AttnDropKey drop,
                   long long* __restrict__ trace, int ablate) {
    using C = TcbCfg<D>;
    constexpr int TCB_SM_WARPS = SW, TCB_SM_THREADS = SW * 32;
    constexpr int TCB_CPT = 256 / SW;                      // key columns per softmax thread and half (32 or 64)
    static_assert(SW == 8 || (SW == 4 && C::CPS == 2), "softmax warps per CTA");
    constexpr int RB = C::RB, TILE = C::TILE, PBYTES = C::PBYTES, NBUF_Q = C::NBUF_Q, NBUF_P = C::NBUF_P;
    constexpr uint32_t LT = umma_layout_for_row_bytes(RB); // swizzle mode of the Q/K/V/dO tiles
    // dQ is double buffered (the drain of tile it - 1 runs under the MMAs of tile it)
    constexpr uint32_t COL_S = 0, COL_DP = 64, COL_DQ = 128, COL_DK = 128 + 2 * D, COL_DV = 128 + 3 * D;
    static_assert(COL_DV + D <= C::TMEM_COLS, "TMEM budget");

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 32-bit shared-space addresses throughout (no generic -> shared conversions in the loops)
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sK = smem;
    const uint32_t sV = sK + TILE;
    const uint32_t sQ = sV + TILE;                  // [NBUF_Q][TILE]
    const uint32_t sdO = sQ + NBUF_Q * TILE;        // [NBUF_Q][TILE]
    const uint32_t sP = sdO + NBUF_Q * TILE;        // [NBUF_P][PBYTES]
    const uint32_t sdS = sP + NBUF_P * PBYTES;      // [NBUF_P][PBYTES]
    const uint32_t bars = sdS + NBUF_P * PBYTES;
    const uint32_t bar_kv = bars;                   // K, V landed
    const uint32_t bar_load = bars + 8;             // [3] Q, dO tile landed
    const uint32_t bar_s_full = bars + 32;          // S, dP of a half complete in TMEM
    const uint32_t bar_sdp_free = bars + 40;        // S, dP of a half pulled into registers by every softmax thread
    const uint32_t bar_p_full = bars + 48;          // both halves of P, dS' written to smem
    // one barrier per dQ buffer: a waiter may then lag a whole tile behind without meeting the phase parity again
    const uint32_t bar_dq_full = bars + 56;         // [2] dV, dK, dQ MMAs of the tile complete
    const uint32_t bar_dq_free = bars + 72;         // [2] dQ buffer drained from TMEM
    const uint32_t tmem_slot = bars + 88;

    const int E = H * D;
    const int kb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int k0 = kb * TCB_TILE;
    const int nq = (T + TCB_TILE - 1) / TCB_TILE;
    const int ntiles = nq - kb;                            // query tiles kb .. nq-1
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    // Diagnostic timeline (cb200_set_attention_trace): lane 0 of every warp of CTA (0, 0, 0) appends (event << 40 | clock)
    // words to its own 512-entry region; nullptr in normal runs (one predictable branch per event).
    long long* tr = (TRACED && trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0) ? trace + warp * 512 : nullptr;
    int tr_n = 0;
    auto TR = [&](int ev) {
        if (TRACED && tr != nullptr && tr_n < 511) tr[++tr_n] = (static_cast<long long>(ev) << 40) | (clock64() & 0xFFFFFFFFFFll);
    };
    TR(1);

    if (warp == TCB_SM_WARPS && lane == 0) {
        tma_prefetch_desc(&tm_qkv);
        tma_prefetch_desc(&tm_do);
        mbar_init_a(bar_kv, 1);
        for (int i = 0; i < 3; ++i) mbar_init_a(bar_load + 8 * i, 1);
        mbar_init_a(bar_s_full, 1);
        mbar_init_a(bar_sdp_free, TCB_SM_THREADS);
        mbar_init_a(bar_p_full, 2 * TCB_SM_THREADS);
        for (int i = 0; i < 2; ++i) {
            mbar_init_a(bar_dq_full + 8 * i, 3);      // one commit per accumulator chain (dV, dK, dQ)
            mbar_init_a(bar_dq_free + 8 * i, TCB_SM_THREADS);
        }
        mbar_fence_init();
        // the first loads do not need TMEM: issue them before the CTA-wide barrier (the allocation may have to wait for
        // the previous CTA of this SM slot to release its columns)
        mbar_expect_tx_a(bar_kv, 2 * TILE);
        tma_load_2d_a(sK, &tm_qkv, bar_kv, H * D + h * D, b * T + k0);
        tma_load_2d_a(sV, &tm_qkv, bar_kv, 2 * H * D + h * D, b * T + k0);
        for (int t = 0; t < (NBUF_Q == 3 ? 2 : 1) && t < ntiles; ++t) {
            mbar_expect_tx_a(bar_load + 8 * t, 2 * TILE);
            tma_load_2d_a(sQ + t * TILE, &tm_qkv, bar_load + 8 * t, h * D, b * T + (kb + t) * TCB_TILE);
            tma_load_2d_a(sdO + t * TILE, &tm_do, bar_load + 8 * t, h * D, b * T + (kb + t) * TCB_TILE);
        }
    }
    if (warp == TCB_SM_WARPS + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");
    const int row_base = b * T;                            // first row of this sequence in the [B*T, ...] tensors

    if (warp >= TCB_SM_WARPS) {
      if (C::CPS == 2) setmaxnreg_dec<C::CONTROL_REGS>();
      // Two issuing threads: an issuing thread runs alone at one instruction every few cycles, and with 28 small MMAs
      // per tile its instruction stream, not the tensor pipe, sets the pace (tools/cuda/umma_bench.cu).  Descriptors
      // are kept as (lo, hi) words: hi is a constant per operand kind, lo = start address field + constant offsets.
      constexpr uint32_t HI_T = umma_desc_hi(8 * RB, LT);        // Q / K / V / dO tiles (rows of RB bytes), either major
      constexpr uint32_t HI_P = umma_desc_hi(1024, 2u);          // P / dS' tiles (rows of 128 bytes, SWIZZLE_128B)
      if (warp == TCB_SM_WARPS) {
        // ===================== loader + S / dP issuer =====================
        if (elect_one()) {
            constexpr uint32_t IDESC_S = umma_idesc_bf16(128, 64, 0, 0);      // Q K_h^T, dO V_h^T
            auto load_tile = [&](int t) {
                const int buf = t % NBUF_Q;
                const int y = row_base + (kb + t) * TCB_TILE;
                mbar_expect_tx_a(bar_load + 8 * buf, 2 * TILE);
                tma_load_2d_a(sQ + buf * TILE, &tm_qkv, bar_load + 8 * buf, h * D, y);
                tma_load_2d_a(sdO + buf * TILE, &tm_do, bar_load + 8 * buf, h * D, y);
            };
            const uint32_t q_lo = umma_desc_lo(sQ, 16), o_lo = umma_desc_lo(sdO, 16);
            const uint32_t k_lo = umma_desc_lo(sK, 16), v_lo = umma_desc_lo(sV, 16);
            auto issue_s_dp = [&](int t, int half) {
                const int buf = t % NBUF_Q;
                if (half == 0) mbar_wait_a(bar_load + 8 * buf, (t / NBUF_Q) & 1);
                tc_fence_after();
                const uint32_t tq = q_lo + buf * (TILE >> 4), to = o_lo + buf * (TILE >> 4);
                const uint32_t hk = half * (64 * RB >> 4);
                // K-major operands, d_h/16 k-steps of 32 bytes inside the swizzle row
#pragma unroll
                for (int ks = 0; ks < D / 16; ++ks)
                    umma_bf16_w(tmem + COL_S, tq + ks * 2, HI_T, k_lo + hk + ks * 2, HI_T, IDESC_S, ks > 0 ? 1u : 0u);
#pragma unroll
                for (int ks = 0; ks < D / 16; ++ks)
                    umma_bf16_w(tmem + COL_DP, to + ks * 2, HI_T, v_lo + hk + ks * 2, HI_T, IDESC_S, ks > 0 ? 1u : 0u);
                umma_commit_a(bar_s_full);
            };
            mbar_wait_a(bar_kv, 0);                        // (K, V and the first Q / dO tiles were requested in the prologue)
            TR(2);
            issue_s_dp(0, 0);
            TR(3);
            for (int it = 0; it < ntiles; ++it) {
                mbar_wait_a(bar_sdp_free, 0);              // half 0 of tile it is in registers: TMEM columns reusable
                TR(10);
                issue_s_dp(it, 1);
                TR(11);
                // the ring slot of tile it + NBUF_Q - 1 was last read by the MMAs of tile it - 1
                if (it >= 1) mbar_wait_a(bar_dq_full + 8 * ((it - 1) & 1), ((it - 1) >> 1) & 1);
                TR(12);
                if (it + NBUF_Q - 1 < ntiles) load_tile(it + NBUF_Q - 1);
                if (it + 1 < ntiles) {
                    mbar_wait_a(bar_sdp_free, 1);          // half 1 of tile it pulled
                    TR(13);
                    issue_s_dp(it + 1, 0);
                    TR(14);
                }
            }
        }
      } else {
        // ===================== dV / dK / dQ issuers: one warp per accumulator chain =====================
        // (8 MMAs each per tile; bar_dq_full counts the three commits)
        if (elect_one()) {
            constexpr uint32_t IDESC_T = umma_idesc_bf16(128, D, 1, 1);       // P^T dO, dS^T Q
            constexpr uint32_t IDESC_Q = umma_idesc_bf16(128, D, 0, 1);       // dS K
            const int chain = warp - TCB_SM_WARPS - 1;                         // 0: dV, 1: dK, 2: dQ
            // A: P^T / dS'^T read MN-major (M = keys, 16384 bytes between the key halves) or dS' K-major
            const uint32_t a_lo = chain == 0 ? umma_desc_lo(sP, 16384) : chain == 1 ? umma_desc_lo(sdS, 16384) : umma_desc_lo(sdS, 16);
            // B: dO / Q / K tile read MN-major
            const uint32_t b_lo = umma_desc_lo(chain == 0 ? sdO : chain == 1 ? sQ : sK, 128 * RB);
            const uint32_t dcol = tmem + (chain == 0 ? COL_DV : chain == 1 ? COL_DK : COL_DQ);
            mbar_wait_a(bar_kv, 0);
            for (int it = 0; it < ntiles; ++it) {
                const int buf = it % NBUF_Q;
                mbar_wait_a(bar_load + 8 * buf, (it / NBUF_Q) & 1);            // (long complete: S / dP of the tile used it)
                mbar_wait_a(bar_p_full, it & 1);
                TR(20);
                if (chain == 2 && it >= 2) mbar_wait_a(bar_dq_free + 8 * (it & 1), ((it - 2) >> 1) & 1);  // tile it - 2 drained from this dQ buffer
                tc_fence_after();
                TR(21);
                const uint32_t pb = (it % NBUF_P) * (PBYTES >> 4);
                if (ablate & (1 << chain)) {
                    // timing-only ablation (CB200_BWD_ABLATE: bit 0 dV, 1 dK, 2 dQ): the chain's MMAs are not issued
                } else if (chain < 2) {
                    const uint32_t tb = b_lo + buf * (TILE >> 4);
                    const uint32_t acc = it > 0 ? 1u : 0u;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        umma_bf16_w(dcol, a_lo + pb + ks * (2048 >> 4), HI_P, tb + ks * (16 * RB >> 4), HI_T, IDESC_T, ks > 0 ? 1u : acc);
                } else {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        umma_bf16_w(dcol + (it & 1) * D, a_lo + pb + (ks >> 2) * (16384 >> 4) + (ks & 3) * 2, HI_P,
                                    b_lo + ks * (16 * RB >> 4), HI_T, IDESC_Q, ks > 0 ? 1u : 0u);
                }
                umma_commit_a(bar_dq_full + 8 * (it & 1));
                TR(22);
            }
        }
      }
    } else {
        // ===================== softmax warps =====================
        if (C::CPS == 2) setmaxnreg_inc<(SW == 4) ? C::SOFTMAX_REGS4 : C::SOFTMAX_REGS>();
        const int quad = warp & 3;                    // TMEM lane quadrant = 32-row band of the tile
        const int cq = warp >> 2;                     // which 32-column slice of a 64-key half
        const int r = quad * 32 + lane;               // row inside the tile
        const uint32_t t_lane = tmem + (static_cast<uint32_t>(quad * 32) << 16);
        const float ks_scale = DROP ? drop.keep_scale : 1.0f;
        const float inv_ks = 1.0f / ks_scale;
        const float* glse = lse + (static_cast<size_t>(b) * H + h) * T;
        const float* gdelta = delta + (static_cast<size_t>(b) * H + h) * T;
        float* dqb = dq_acc + static_cast<size_t>(b) * T * E + h * D;
        // dropout: the row's stream of this 128-key block (common.cuh); this thread starts at pair 32 half + 16 cq
        const float thr = __uint_as_float(drop.thr_bits);
        const uint32_t drop_base = DROP ? attn_drop_base(drop, b * H + h) : 0u;
        const uint32_t drop_jump0 = cq ? mcg_mul_pow(16) : 1u;
        const uint32_t drop_jump1 = cq ? mcg_mul_pow(48) : mcg_mul_pow(32);
        const uint64_t sc2 = f2_pack(scale_log2, scale_log2);
        const float dq_scale = scale * ks_scale;

        auto drain_dq = [&](int t) {                  // tile t's dQ: the two warps of a row band take half the columns each
            constexpr int DC = D / (SW / 4);          // columns per thread: the SW / 4 warps of a row band share the row
            const int row_g = (kb + t) * TCB_TILE + r;
            const int buf = t & 1;
            mbar_wait_a(bar_dq_full + 8 * buf, (t >> 1) & 1);
            tc_fence_after();
            uint32_t v[DC];
            if (DC == 8) tmem_ld8(t_lane + COL_DQ + buf * D + cq * DC, reinterpret_cast<uint32_t(&)[8]>(v));
            else if (DC == 16) tmem_ld16(t_lane + COL_DQ + buf * D + cq * DC, reinterpret_cast<uint32_t(&)[16]>(v));
            else tmem_ld32(t_lane + COL_DQ + buf * D + cq * DC, reinterpret_cast<uint32_t(&)[32]>(v));
            tmem_ld_wait();
            if (row_g < T) {
                float* dst = dqb + static_cast<size_t>(row_g) * E + cq * DC;
#pragma unroll
                for (int j = 0; j < DC; j += 4)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j),
                                 "f"(__uint_as_float(v[j]) * dq_scale), "f"(__uint_as_float(v[j + 1]) * dq_scale),
                                 "f"(__uint_as_float(v[j + 2]) * dq_scale), "f"(__uint_as_float(v[j + 3]) * dq_scale)
                                 : "memory");
            }
            tc_fence_before();
            mbar_arrive_a(bar_dq_free + 8 * buf);
        };

        // per-row softmax statistics are fetched one tile ahead (a global-memory latency per tile otherwise)
        float lse_next = (kb * TCB_TILE + r < T) ? glse[kb * TCB_TILE + r] : INFINITY;   // +inf => P = 0 past the end
        float dl_next = (kb * TCB_TILE + r < T) ? gdelta[kb * TCB_TILE + r] : 0.f;      // raw: scaled when it is used, a tile later
        for (int it = 0; it < ntiles; ++it) {
            const int row_g = (kb + it) * TCB_TILE + r;            // global query row
            const float lse_r = lse_next, dl_r = dl_next * inv_ks;
            {
                const int row_n = row_g + TCB_TILE;
                const bool ok = (it + 1 < ntiles) && row_n < T;
                lse_next = ok ? glse[row_n] : INFINITY;
                dl_next = ok ? gdelta[row_n] : 0.f;
            }
            const bool diagonal = (it == 0);
            const uint32_t bp_a = sP + (it % NBUF_P) * PBYTES;
            const uint32_t bs_a = sdS + (it % NBUF_P) * PBYTES;
            const uint32_t x0 = DROP ? attn_row_seed(drop_base, static_cast<uint32_t>(row_g), static_cast<uint32_t>(kb)) : 0u;
            const uint64_t nlse2 = f2_pack(-lse_r, -lse_r), ndl2 = f2_pack(-dl_r, -dl_r);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int col_base = half * 64 + cq * TCB_CPT;     // first key column of this thread's slice (inside the tile)
                mbar_wait_a(bar_s_full, half);
                tc_fence_after();
                TR(30 + half);
                // The thread's 32 columns of S and dP are pulled out of TMEM at once and the columns are released
                // immediately: the control thread issues S / dP of the next half under this half's arithmetic.
                uint32_t sv_all[TCB_CPT], dv_all[TCB_CPT];
                const bool skip_all = diagonal && col_base > quad * 32 + 31;   // every key of the slice is above the band
                if (!skip_all) {
#pragma unroll
                    for (int c32 = 0; c32 < TCB_CPT; c32 += 32) {
                        tmem_ld32(t_lane + COL_S + cq * TCB_CPT + c32, *reinterpret_cast<uint32_t(*)[32]>(&sv_all[c32]));
                        tmem_ld32(t_lane + COL_DP + cq * TCB_CPT + c32, *reinterpret_cast<uint32_t(*)[32]>(&dv_all[c32]));
                    }
                    tmem_ld_wait();
                }
                tc_fence_before();
                mbar_arrive_a(bar_sdp_free);
                TR(32 + half);
                uint32_t x = x0 * (half ? drop_jump1 : drop_jump0);
                const uint32_t row_off = static_cast<uint32_t>(half * 16384 + r * 128);
                // 16 columns at a time keep the live register set small
#pragma unroll
                for (int ch = 0; ch < TCB_CPT / 16; ++ch) {
                    const int col0 = col_base + ch * 16;           // first key column of this chunk (inside the tile)
                    uint32_t pk[8], dk_[8];                        // packed bf16 pairs of P (dropped) and dS'
                    if (diagonal && col0 > quad * 32 + 31) {
                        // every key of this chunk is above every row of this band (and so is every later chunk of the thread)
#pragma unroll
                        for (int i = 0; i < 8; ++i) { pk[i] = 0u; dk_[i] = 0u; }
                    } else {
                        const uint32_t* sv = sv_all + 16 * ch;
                        const uint32_t* dv_ = dv_all + 16 * ch;
                        const bool partial = diagonal && (col0 + 15 > quad * 32);
                        // two copies of the pair loop: only chunks that straddle the diagonal pay for the mask.
                        // Per pair: FFMA2, 2 MUFU, [2 IMAD, 2 FSET, FMUL2], FFMA2 / FADD2, FMUL2, 2 F2FP.
                        auto pairs = [&](auto masked) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const uint64_t t2 = f2_fma(f2_pack(sv[2 * q], sv[2 * q + 1]), sc2, nlse2);
                                float t0, t1;
                                f2_unpack(t2, t0, t1);
                                float p0 = fast_exp2(t0), p1 = fast_exp2(t1);
                                if (decltype(masked)::value) {
                                    if (col0 + 2 * q > r) p0 = 0.f;
                                    if (col0 + 2 * q + 1 > r) p1 = 0.f;
                                }
                                const uint64_t p2 = f2_pack(p0, p1);
                                const uint64_t dp2 = f2_pack(dv_[2 * q], dv_[2 * q + 1]);
                                uint64_t u2, pm2;
                                if (DROP) {
                                    float m0, m1;
                                    attn_drop_pair(x, thr, m0, m1);
                                    const uint64_t m2 = f2_pack(m0, m1);
                                    u2 = f2_fma(m2, dp2, ndl2);        // M.dP - delta/ks
                                    pm2 = f2_mul(p2, m2);
                                } else {
                                    u2 = f2_add(dp2, ndl2);
                                    pm2 = p2;
                                }
                                const uint64_t ds2 = f2_mul(p2, u2);
                                float a0, a1, b0, b1;
                                f2_unpack(pm2, a0, a1);
                                f2_unpack(ds2, b0, b1);
                                pk[q] = pack_bf16(a0, a1);
                                dk_[q] = pack_bf16(b0, b1);
                            }
                        };
                        if (partial) pairs(std::true_type{});
                        else         pairs(std::false_type{});
                    }
                    // single P / dS' buffer: the MMAs of tile it - 1 must have read it before it is overwritten
                    if (NBUF_P == 1 && half == 0 && ch == 0 && it >= 1) {
                        TR(34);
                        mbar_wait_a(bar_dq_full + 8 * ((it - 1) & 1), ((it - 1) >> 1) & 1);
                        TR(35);
                    }
                    // two 16-byte pieces of this row per tensor; piece index XOR (row & 7) = SWIZZLE_128B
#pragma unroll
                    for (int c2 = 0; c2 < 2; ++c2) {
                        const uint32_t off = row_off + (((cq * 4 + ch * 2 + c2) ^ (r & 7)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(bp_a + off), "r"(pk[4 * c2]), "r"(pk[4 * c2 + 1]),
                                     "r"(pk[4 * c2 + 2]), "r"(pk[4 * c2 + 3]) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(bs_a + off), "r"(dk_[4 * c2]), "r"(dk_[4 * c2 + 1]),
                                     "r"(dk_[4 * c2 + 2]), "r"(dk_[4 * c2 + 3]) : "memory");
                    }
                }
                fence_proxy_async_smem();
                mbar_arrive_a(bar_p_full);
                TR(36 + half);
            }
            // dQ of the previous tile: its MMAs were issued a whole tile of softmax work ago
            if (it >= 1) drain_dq(it - 1);
            TR(38);
        }
        // ---- last dQ tile, then dK / dV of this key block (thread = key row) ----
        drain_dq(ntiles - 1);
        if (cq == 0) {
            const int key = k0 + r;
            const int ld = 3 * E;
            __nv_bfloat16* dkp = dqkv + (static_cast<size_t>(row_base) + key) * ld + E + h * D;
            __nv_bfloat16* dvp = dqkv + (static_cast<size_t>(row_base) + key) * ld + 2 * E + h * D;
            const float dk_scale = scale * ks_scale;
#pragma unroll
            for (int c0 = 0; c0 < D; c0 += 16) {
                uint32_t vk[16], vv[16];
                tmem_ld16(t_lane + COL_DK + c0, vk);
                tmem_ld16(t_lane + COL_DV + c0, vv);
                tmem_ld_wait();
                if (key < T) {
                    uint32_t ok[8], ov[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        ok[j] = pack_bf16(__uint_as_float(vk[2 * j]) * dk_scale, __uint_as_float(vk[2 * j + 1]) * dk_scale);
                        ov[j] = pack_bf16(__uint_as_float(vv[2 * j]) * ks_scale, __uint_as_float(vv[2 * j + 1]) * ks_scale);
                    }
                    reinterpret_cast<uint4*>(dkp + c0)[0] = make_uint4(ok[0], ok[1], ok[2], ok[3]);
                    reinterpret_cast<uint4*>(dkp + c0)[1] = make_uint4(ok[4], ok[5], ok[6], ok[7]);
                    reinterpret_cast<uint4*>(dvp + c0)[0] = make_uint4(ov[0], ov[1], ov[2], ov[3]);
                    reinterpret_cast<uint4*>(dvp + c0)[1] = make_uint4(ov[4], ov[5], ov[6], ov[7]);
                }
            }
        }
    }

    TR(99);
    if (TRACED && tr != nullptr) tr[0] = tr_n;
    tc_fence_before();
    __syncthreads();
    if (warp == TCB_SM_WARPS + 1) {
        tc_fence_after();
        tmem_dealloc<C::TMEM_COLS>(tmem);
    }
}

template <int D, bool DROP, int SW = 8, bool TRACED = false>
static int launch_bwd_tc(const __nv_bfloat16* qkv, const __nv_bfloat16* dout, const float* lse, const float* delta,
                         float* dq_acc, __nv_bfloat16* dqkv, int B, int T, int H, float scale, const AttnDropKey& key,
                         cudaStream_t s) {
    constexpr int RB = 2 * D;
    constexpr size_t smem = TcbCfg<D>::SMEM;
    static_assert(TcbCfg<D>::CPS * (smem + 1024) <= 233472, "shared memory budget");
    const int E = H * D;
    CUtensorMap tm_qkv, tm_do;
    int rc = make_tmap_bf16_sw(&tm_qkv, qkv, 3 * E, static_cast<uint64_t>(B) * T, 3 * E, D, TCB_TILE, RB);
    if (rc) return rc;
    rc = make_tmap_bf16_sw(&tm_do, dout, E, static_cast<uint64_t>(B) * T, E, D, TCB_TILE, RB);
    if (rc) return rc;
    auto kernel = attn_bwd_tc_kernel<D, DROP, SW, TRACED>;
    static const int ablate = getenv("CB200_BWD_ABLATE") ? atoi(getenv("CB200_BWD_ABLATE")) : 0;   // diagnostic, results wrong
    static bool configured = false;
    if (!configured) {
        CB200_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    // longest walks first: x = 0 is the key block with every query tile below it
    dim3 grid((T + TCB_TILE - 1) / TCB_TILE, H, B);
    kernel<<<grid, (SW + 4) * 32, smem, s>>>(tm_qkv, tm_do, lse, delta, dq_acc, dqkv, T, H, scale,
                                            scale * 1.4426950408889634f, key, g_attention_trace, ablate);
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

// The tcgen05 main kernel of attention_bwd (delta and the dq store stay in attention.cu).
int attention_bwd_tc_main(const __nv_bfloat16* qkv, const __nv_bfloat16* dout, const float* lse, const float* delta,
                          float* dq_acc, __nv_bfloat16* dqkv, int B, int T, int H, int D, float scale,
                          const AttnDropKey& key, cudaStream_t s) {
    const bool dropping = key.thr_bits != 0;
    static const int sw4 = getenv("CB200_BWD_SW4") ? atoi(getenv("CB200_BWD_SW4")) : 0;   // A/B: 4 softmax warps per CTA (d_h 16)
    if (D == 16 && g_attention_trace != nullptr)          // diagnostic: the default shape with the event timeline
        return dropping ? launch_bwd_tc<16, true, 8, true>(qkv, dout, lse, delta, dq_acc, dqkv, B, T, H, scale, key, s)
                        : launch_bwd_tc<16, false, 8, true>(qkv, dout, lse, delta, dq_acc, dqkv, B, T, H, scale, key, s);
    if (D == 16 && sw4)
        return dropping ? launch_bwd_tc<16, true, 4>(qkv, dout, lse, delta, dq_acc, dqkv, B, T, H, scale, key, s)
                        : launch_bwd_tc<16, false, 4>(qkv, dout, lse, delta, dq_acc, dqkv, B, T, H, scale, key, s);
    switch (D) {
        case 16: return dropping ? launch_bwd_tc<16, true>(qkv, dout, lse, delta, dq_acc, dqkv, B, T, H, scale, key, s)
                                 : launch_bwd_tc<16, false>(qkv, dout, lse, delta, dq_acc, dqkv, B, T, H, scale, key, s);
        case 32: return dropping ? launch_bwd_tc<32, true>(qkv, dout, lse, delta, dq_acc, dqkv, B, T, H, scale, key, s)
                                 : launch_bwd_tc<32, false>(qkv, dout, lse, delta, dq_acc, dqkv, B, T, H, scale, key, s);
        case 64: return dropping ? launch_bwd_tc<64, true>(qkv, dout, lse, delta, dq_acc, dqkv, B, T, H, scale, key, s)
                                 : launch_bwd_tc<64, false>(qkv, dout, lse, delta, dq_acc, dqkv, B, T, H, scale, key, s);
        default: break;
    }
    set_error("attention head size %d is not supported (16, 32 or 64)", D);
    return -1;
}

}  // namespace cb200
