// Attention backward on the 5th-generation tensor cores (tcgen05 / TMEM).
//
// One CTA owns 128 keys of one (batch, head) and walks the 128-row query tiles
// at or below the diagonal.  Per tile:
//   S  = Q K^T, dP = dO V^T          tcgen05.mma, M 128 x N 128, accumulators in TMEM
//   softmax warps (thread = query row, 32 columns each): P = exp2(S c - lse),
//     causal mask, dropout, dS' = P (M.dP - delta/ks) -> bf16 -> swizzled smem
//   dV += P^T dO, dK += dS'^T Q      the same smem tile read through an MN-major
//   dQ  = dS' K                       descriptor (transposed) and a K-major one
// so no score element is ever transposed or re-loaded by a CUDA core: compared
// with the mma.sync kernel (attention.cu) the movmatrix / ldmatrix / HMMA issue
// slots disappear and the element work is the only thing the SM issues.
// dK, dV accumulate in TMEM over the whole walk; dQ tiles are drained from TMEM
// by the thread that owns the row and reduced into the fp32 dq buffer.
//
// Q, K, V, dO tiles are rows of d_h bf16 (32 / 64 / 128 bytes) loaded by TMA
// with the swizzle whose span is one row, consumed K-major (S, dP) and MN-major
// (dV, dK, dQ right operands).  Reference semantics: see attention.cu.
#include "attention.h"
#include "gemm.h"

#include <type_traits>

namespace cb200 {

constexpr int TCB_SM_WARPS = 16;                       // softmax warps: 4 row bands x 4 column quarters
constexpr int TCB_THREADS = (TCB_SM_WARPS + 4) * 32;   // + one warpgroup: control warp, TMEM allocator warp, 2 idle
// registers after setmaxnreg: the CTA is launched with 96 registers per thread (640 threads); the control warpgroup
// keeps 32 and what it releases, 128 * 64, lets the 16 softmax warps grow by 16 each: they hold their 32 columns of S
// and of dP (64 registers) while the next tile's MMAs already run.  (setmaxnreg.inc can only take what the CTA's own
// warps released: 512 * (112 - 96) == 128 * (96 - 32).)
constexpr int TCB_LAUNCH_REGS = 96, TCB_SOFTMAX_REGS = 112, TCB_CONTROL_REGS = 32;
static_assert(TCB_SM_WARPS * 32 * (TCB_SOFTMAX_REGS - TCB_LAUNCH_REGS) <= 128 * (TCB_LAUNCH_REGS - TCB_CONTROL_REGS), "register pool");
constexpr int TCB_CPT = 128 / (TCB_SM_WARPS / 4);      // key columns per softmax thread (32)
constexpr int TCB_TILE = 128;                          // query rows per tile = keys per CTA

template <int D>
struct TcbCfg {
    static constexpr int RB = 2 * D;                       // bytes per row of a Q/K/V/dO tile
    static constexpr int TILE = TCB_TILE * RB;             // 4 / 8 / 16 KB
    static constexpr int PBYTES = 2 * TCB_TILE * 128;      // P or dS': [2 column halves][128 rows][128 B]
    static constexpr int NBUF_Q = (D <= 32) ? 3 : 2;       // Q / dO tile ring
    static constexpr int NBUF_P = (D <= 32) ? 2 : 1;       // P / dS' buffers
    static constexpr size_t SMEM = 2 * TILE + 2 * NBUF_Q * TILE + 2 * NBUF_P * PBYTES + 256 + 1024;
};

// Software pipeline (tile index it):
//   control thread:  S/dP(it+1) is issued as soon as the softmax warps have pulled S/dP(it) out of TMEM
//                    (bar_sdp_free), i.e. it runs under the softmax arithmetic of tile it; dV/dK/dQ(it) follow
//                    when P/dS'(it) are in shared memory (bar_p_full).
//   softmax warps:   wait S/dP(it) -> tcgen05.ld -> release TMEM -> arithmetic -> smem -> bar_p_full;
//                    dQ(it-1) is drained at the top of iteration it, when its MMAs have long finished.
template <int D, bool DROP>
__global__ void __launch_bounds__(TCB_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                   const float* __restrict__ lse, const float* __restrict__ delta, float* __restrict__ dq_acc,
                   __nv_bfloat16* __restrict__ dqkv, int T, int H, float scale, float scale_log2, AttnDropKey drop) {
    using C = TcbCfg<D>;
    constexpr int RB = C::RB, TILE = C::TILE, PBYTES = C::PBYTES, NBUF_Q = C::NBUF_Q, NBUF_P = C::NBUF_P;
    constexpr uint32_t LT = umma_layout_for_row_bytes(RB); // swizzle mode of the Q/K/V/dO tiles
    // dQ is double buffered (the drain of tile it - 1 runs under the MMAs of tile it)
    constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DQ = 256, COL_DK = 256 + 2 * D, COL_DV = 256 + 3 * D;
    static_assert(COL_DV + D <= 512, "TMEM budget");

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
    const uint32_t bar_s_full = bars + 32;          // S, dP complete in TMEM
    const uint32_t bar_sdp_free = bars + 40;        // S, dP pulled into registers by every softmax thread
    const uint32_t bar_p_full = bars + 48;          // P, dS' written to smem
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

    if (warp == TCB_SM_WARPS && lane == 0) {
        tma_prefetch_desc(&tm_qkv);
        tma_prefetch_desc(&tm_do);
        mbar_init_a(bar_kv, 1);
        for (int i = 0; i < 3; ++i) mbar_init_a(bar_load + 8 * i, 1);
        mbar_init_a(bar_s_full, 1);
        mbar_init_a(bar_sdp_free, TCB_SM_WARPS * 32);
        mbar_init_a(bar_p_full, TCB_SM_WARPS * 32);
        for (int i = 0; i < 2; ++i) {
            mbar_init_a(bar_dq_full + 8 * i, 1);
            mbar_init_a(bar_dq_free + 8 * i, TCB_SM_WARPS * 32);
        }
        mbar_fence_init();
    }
    if (warp == TCB_SM_WARPS + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");
    const int row_base = b * T;                            // first row of this sequence in the [B*T, ...] tensors

    if (warp >= TCB_SM_WARPS) {
      setmaxnreg_dec<TCB_CONTROL_REGS>();
      if (warp == TCB_SM_WARPS) {
        // ===================== control warp: TMA producer + MMA issuer =====================
        if (elect_one()) {
            constexpr uint32_t IDESC_S = umma_idesc_bf16(128, 128, 0, 0);     // Q K^T, dO V^T
            constexpr uint32_t IDESC_T = umma_idesc_bf16(128, D, 1, 1);       // P^T dO, dS^T Q
            constexpr uint32_t IDESC_Q = umma_idesc_bf16(128, D, 0, 1);       // dS K
            const uint32_t aK = sK, aV = sV;
            auto load_tile = [&](int t) {
                const int buf = t % NBUF_Q;
                const int y = row_base + (kb + t) * TCB_TILE;
                mbar_expect_tx_a(bar_load + 8 * buf, 2 * TILE);
                tma_load_2d_a(sQ + buf * TILE, &tm_qkv, bar_load + 8 * buf, h * D, y);
                tma_load_2d_a(sdO + buf * TILE, &tm_do, bar_load + 8 * buf, h * D, y);
            };
            auto issue_s_dp = [&](int t) {
                const int buf = t % NBUF_Q;
                mbar_wait_a(bar_load + 8 * buf, (t / NBUF_Q) & 1);
                tc_fence_after();
                const uint32_t aQ = sQ + buf * TILE, aO = sdO + buf * TILE;
                // K-major operands, d_h/16 k-steps of 32 bytes inside the swizzle row
#pragma unroll
                for (int ks = 0; ks < D / 16; ++ks)
                    umma_bf16(tmem + COL_S, umma_smem_desc(aQ + ks * 32, 16, 8 * RB, LT),
                              umma_smem_desc(aK + ks * 32, 16, 8 * RB, LT), IDESC_S, ks > 0 ? 1u : 0u);
#pragma unroll
                for (int ks = 0; ks < D / 16; ++ks)
                    umma_bf16(tmem + COL_DP, umma_smem_desc(aO + ks * 32, 16, 8 * RB, LT),
                              umma_smem_desc(aV + ks * 32, 16, 8 * RB, LT), IDESC_S, ks > 0 ? 1u : 0u);
                umma_commit_a(bar_s_full);
            };
            mbar_expect_tx_a(bar_kv, 2 * TILE);
            tma_load_2d_a(sK, &tm_qkv, bar_kv, E + h * D, row_base + k0);
            tma_load_2d_a(sV, &tm_qkv, bar_kv, 2 * E + h * D, row_base + k0);
            load_tile(0);
            if (NBUF_Q == 3 && ntiles > 1) load_tile(1);
            mbar_wait_a(bar_kv, 0);
            issue_s_dp(0);
            for (int it = 0; it < ntiles; ++it) {
                // the ring slot of tile it + NBUF_Q - 1 was last read by the MMAs of tile it - 1
                if (it >= 1) mbar_wait_a(bar_dq_full + 8 * ((it - 1) & 1), ((it - 1) >> 1) & 1);
                if (it + NBUF_Q - 1 < ntiles) load_tile(it + NBUF_Q - 1);
                if (it + 1 < ntiles) {
                    mbar_wait_a(bar_sdp_free, it & 1);     // S/dP(it) are in registers: TMEM columns reusable
                    tc_fence_after();
                    issue_s_dp(it + 1);
                }
                mbar_wait_a(bar_p_full, it & 1);
                tc_fence_after();
                if (it >= 2) {
                    mbar_wait_a(bar_dq_free + 8 * (it & 1), ((it - 2) >> 1) & 1);  // tile it - 2 drained from this dQ buffer
                    tc_fence_after();
                }
                const int buf = it % NBUF_Q;
                const uint32_t aQ = sQ + buf * TILE, aO = sdO + buf * TILE;
                const uint32_t aP = sP + (it % NBUF_P) * PBYTES, aS = sdS + (it % NBUF_P) * PBYTES;
                // dV += P^T dO, dK += dS'^T Q : A = smem tile read MN-major (M = keys), K = 128 query rows
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    umma_bf16(tmem + COL_DV, umma_smem_desc(aP + ks * 2048, 16384, 1024, 2u),
                              umma_smem_desc(aO + ks * 16 * RB, 128 * RB, 8 * RB, LT), IDESC_T, (it > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    umma_bf16(tmem + COL_DK, umma_smem_desc(aS + ks * 2048, 16384, 1024, 2u),
                              umma_smem_desc(aQ + ks * 16 * RB, 128 * RB, 8 * RB, LT), IDESC_T, (it > 0 || ks > 0) ? 1u : 0u);
                // dQ = dS' K : A K-major (two 64-key halves of 16 KB), B = K tile MN-major
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    umma_bf16(tmem + COL_DQ + (it & 1) * D, umma_smem_desc(aS + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024, 2u),
                              umma_smem_desc(aK + ks * 16 * RB, 128 * RB, 8 * RB, LT), IDESC_Q, ks > 0 ? 1u : 0u);
                umma_commit_a(bar_dq_full + 8 * (it & 1));
            }
        }
      }
    } else {
        // ===================== softmax warps =====================
        setmaxnreg_inc<TCB_SOFTMAX_REGS>();
        const int quad = warp & 3;                    // TMEM lane quadrant = 32-row band of the tile
        const int cq = warp >> 2;                     // which TCB_CPT-column slice of the 128 keys
        const int r = quad * 32 + lane;               // row inside the tile
        const uint32_t t_lane = tmem + (static_cast<uint32_t>(quad * 32) << 16);
        const float ks_scale = DROP ? drop.keep_scale : 1.0f;
        const float inv_ks = 1.0f / ks_scale;
        const float* glse = lse + (static_cast<size_t>(b) * H + h) * T;
        const float* gdelta = delta + (static_cast<size_t>(b) * H + h) * T;
        float* dqb = dq_acc + static_cast<size_t>(b) * T * E + h * D;
        // dropout: the row's stream of this 128-key block (common.cuh); this thread starts at pair 16 * cq
        const float thr = __uint_as_float(drop.thr_bits);
        const uint32_t drop_base = DROP ? attn_drop_base(drop, b * H + h) : 0u;
        const uint32_t drop_jump = (cq == 0) ? 1u : (cq == 1) ? mcg_mul_pow(16) : (cq == 2) ? mcg_mul_pow(32) : mcg_mul_pow(48);
        const uint64_t sc2 = f2_pack(scale_log2, scale_log2);
        const float dq_scale = scale * ks_scale;

        auto drain_dq = [&](int t) {                  // tile t's dQ: every warp reduces a quarter of the head's columns
            constexpr int DC = D / 4;
            const int row_g = (kb + t) * TCB_TILE + r;
            const int buf = t & 1;
            mbar_wait_a(bar_dq_full + 8 * buf, (t >> 1) & 1);
            tc_fence_after();
            uint32_t v[DC];
            if (DC == 4) tmem_ld4(t_lane + COL_DQ + buf * D + cq * DC, reinterpret_cast<uint32_t(&)[4]>(v));
            else if (DC == 8) tmem_ld8(t_lane + COL_DQ + buf * D + cq * DC, reinterpret_cast<uint32_t(&)[8]>(v));
            else tmem_ld16(t_lane + COL_DQ + buf * D + cq * DC, reinterpret_cast<uint32_t(&)[16]>(v));
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
            mbar_wait_a(bar_s_full, it & 1);
            tc_fence_after();
            const bool diagonal = (it == 0);
            const uint32_t bp_a = sP + (it % NBUF_P) * PBYTES;
            const uint32_t bs_a = sdS + (it % NBUF_P) * PBYTES;
            uint32_t x = 0;
            if (DROP) x = attn_row_seed(drop_base, static_cast<uint32_t>(row_g), static_cast<uint32_t>(kb)) * drop_jump;
            const uint64_t nlse2 = f2_pack(-lse_r, -lse_r), ndl2 = f2_pack(-dl_r, -dl_r);
            // The thread's 32 columns of S and dP are pulled out of TMEM at once and the columns are released
            // immediately: the control thread issues S / dP of the next tile under this tile's arithmetic.
            uint32_t sv_all[TCB_CPT], dv_all[TCB_CPT];
            if (!(diagonal && cq > quad)) {                        // else: every key of the slice is above the band
                tmem_ld32(t_lane + COL_S + cq * TCB_CPT, sv_all);
                tmem_ld32(t_lane + COL_DP + cq * TCB_CPT, dv_all);
                tmem_ld_wait();
            }
            tc_fence_before();
            mbar_arrive_a(bar_sdp_free);
#pragma unroll
            for (int ch = 0; ch < TCB_CPT / 16; ++ch) {
                const int col0 = cq * TCB_CPT + ch * 16;           // first key column of this chunk (inside the tile)
                uint32_t pk[8], dk_[8];                            // packed bf16 pairs of P (dropped) and dS'
                if (diagonal && col0 > quad * 32 + 31) {
                    // every key of this chunk is above every row of this band (and so is every later chunk of the thread)
#pragma unroll
                    for (int i = 0; i < 8; ++i) { pk[i] = 0u; dk_[i] = 0u; }
                } else {
                    const uint32_t* sv = sv_all + 16 * ch;
                    const uint32_t* dv_ = dv_all + 16 * ch;
                    const bool partial = diagonal && (col0 + 15 > quad * 32);
                    // two copies of the pair loop: only chunks that straddle the diagonal pay for the mask.
                    // Per pair: FFMA2, 2 MUFU, [IMAD.WIDE, 2 FSET, FMUL2], FFMA2 / FADD2, FMUL2, 2 F2FP.
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
                if (ch == 0 && NBUF_P == 1 && it >= 1) mbar_wait_a(bar_dq_full + 8 * ((it - 1) & 1), ((it - 1) >> 1) & 1);   // single buffer: tile it-1's MMAs must be done
                // two 16-byte chunks of this row per tensor; chunk index XOR (row & 7) = SWIZZLE_128B
                const uint32_t half_off = static_cast<uint32_t>((col0 >> 6) * 16384 + r * 128);
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2) {
                    const uint32_t off = half_off + (((((col0 & 63) >> 3) + c2) ^ (r & 7)) << 4);
                    // explicit shared-space stores with 32-bit addresses (the aligned base pointer went through an
                    // integer cast, so a plain store would be a generic ST with 64-bit address arithmetic)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(bp_a + off), "r"(pk[4 * c2]), "r"(pk[4 * c2 + 1]),
                                 "r"(pk[4 * c2 + 2]), "r"(pk[4 * c2 + 3]) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(bs_a + off), "r"(dk_[4 * c2]), "r"(dk_[4 * c2 + 1]),
                                 "r"(dk_[4 * c2 + 2]), "r"(dk_[4 * c2 + 3]) : "memory");
                }
            }
            fence_proxy_async_smem();
            mbar_arrive_a(bar_p_full);
            // dQ of the previous tile: its MMAs were issued a whole tile of softmax work ago
            if (it >= 1) drain_dq(it - 1);
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

    tc_fence_before();
    __syncthreads();
    if (warp == TCB_SM_WARPS + 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

template <int D, bool DROP>
static int launch_bwd_tc(const __nv_bfloat16* qkv, const __nv_bfloat16* dout, const float* lse, const float* delta,
                         float* dq_acc, __nv_bfloat16* dqkv, int B, int T, int H, float scale, const AttnDropKey& key,
                         cudaStream_t s) {
    constexpr int RB = 2 * D;
    constexpr size_t smem = TcbCfg<D>::SMEM;
    const int E = H * D;
    CUtensorMap tm_qkv, tm_do;
    int rc = make_tmap_bf16_sw(&tm_qkv, qkv, 3 * E, static_cast<uint64_t>(B) * T, 3 * E, D, TCB_TILE, RB);
    if (rc) return rc;
    rc = make_tmap_bf16_sw(&tm_do, dout, E, static_cast<uint64_t>(B) * T, E, D, TCB_TILE, RB);
    if (rc) return rc;
    auto kernel = attn_bwd_tc_kernel<D, DROP>;
    static bool configured = false;
    if (!configured) {
        CB200_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid((T + TCB_TILE - 1) / TCB_TILE, H, B);
    kernel<<<grid, TCB_THREADS, smem, s>>>(tm_qkv, tm_do, lse, delta, dq_acc, dqkv, T, H, scale,
                                            scale * 1.4426950408889634f, key);
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

// The tcgen05 main kernel of attention_bwd (delta and the dq store stay in attention.cu).
int attention_bwd_tc_main(const __nv_bfloat16* qkv, const __nv_bfloat16* dout, const float* lse, const float* delta,
                          float* dq_acc, __nv_bfloat16* dqkv, int B, int T, int H, int D, float scale,
                          const AttnDropKey& key, cudaStream_t s) {
    const bool dropping = key.thr_bits != 0;
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
