// Causal attention forward on the 5th-generation tensor cores (tcgen05 / TMEM), fed by TMA.
//
// Reference semantics (composer/models/transformer.py:331-371): S = q k^T, S *= rsqrt(d_h) (:345-348, before the
// mask), causal mask S*b - 1e4*(1-b) (:351-354; the masked probabilities underflow to exactly 0 in fp32, so they
// are skipped here), softmax (:360), dropout on the probabilities (:361), P v (:367); heads split / merged as
// :373-395, i.e. head h owns columns [h*d_h, (h+1)*d_h) of the q | k | v thirds of c_attn's output.
//
// Persistent CTAs (two per SM) walk work items (one 128-row query tile of one (batch, head)), heaviest first:
//   warp 5 (one lane)  TMA producer: Q tile, then the K / V tiles of the item through a ring (SWIZZLE = row bytes)
//   warp 4 (one lane)  S issuer:     S = Q K^T  (tcgen05.mma, M 128 x N KT, fp32 in TMEM, double buffered), issued
//                                    one tile ahead of the softmax
//   warp 7 (one lane)  P V issuer:   O~ = P V  (M 128 x N d_h) once P is ready
//   warps 0-3          softmax, thread = query row (TMEM lane): tcgen05.ld of the whole row, running max, exp2 with
//                      the scale folded into one FFMA2, row sum, dropout, bf16 P written back INTO THE S BUFFER'S
//                      COLUMNS with tcgen05.st and consumed from there as the A operand (TS form of tcgen05.mma),
//                      so the probabilities never cross shared memory (PSMEM variant: swizzled smem tile, kept for
//                      A/B).  O~ of a tile lands in spare columns of the same buffer and is folded into the thread's
//                      fp32 output row one tile later (o = o*alpha + O~): with d_h <= 64 that is cheaper than
//                      rescaling an accumulator in TMEM and needs no correction warps.
// At d_h = 16 a 128 x 128 tile is ~130 tensor-pipe cycles against 1,024 MUFU cycles (16 ex2 / clk / SM), so the
// budget is instructions per score element: packed fp32 pairs (FFMA2 / FADD2 / FMUL2), a 3-input max, and a dropout
// stream of 1.5 instructions per element (common.cuh) keep the loop at ~5 per element, below the MUFU bound of 8.
#include "attention.h"
#include "gemm.h"

#include <type_traits>

namespace cb200 {

constexpr int TCF_THREADS = 256;          // row split 1; (4 RS + 4) * 32 in general

// KT: keys per tile = the part of a score row a thread holds in registers.  CPS: CTAs per SM.  Four softmax warps
// per CTA means CPS softmax warps per scheduler: KT 128 needs ~180 registers per softmax thread (2 CTAs per SM), KT 64
// about 110 (3 or 4 per SM: more warps to hide the MUFU / TMEM / barrier latencies, twice the per-tile hand-offs).
// RS: threads per score row.  With RS = 2 a row's KT keys are split between two warps of the same TMEM lane quadrant
// (8 softmax warps per CTA, 4 per scheduler with two CTAs per SM): one warp alone reaches ~70 % of the MUFU rate
// (tools/cuda/pipe_bench.cu), so the exponentials only run at full rate while two warps of a scheduler are in their
// arithmetic phase at the same time, which two softmax warps per scheduler rarely are.
template <int D, bool PSMEM, int KT_, int CPS_, int RS_ = 1>
struct TcfCfg {
    static constexpr int KT = KT_, CPS = CPS_, RS = RS_;
    static constexpr int THREADS = (4 * RS_ + 4) * 32;
    static constexpr int RB = 2 * D;                        // bytes per row of a Q / K / V tile
    static constexpr int QTILE = 128 * RB;
    static constexpr int KTILE = KT * RB;
    static constexpr int NKV = (D == 64) ? 3 : 4;           // K / V ring stages
    static constexpr int PBYTES = (KT / 64) * 128 * 128;    // P in shared memory: 64-key halves of [128 rows][128 B]
    static constexpr int XBYTES = (RS_ > 1) ? 3 * RS_ * 128 * 4 : 0;     // row maxima (2 slots, by tile parity) / sums (1 slot) exchanged between the RS threads of a row
    static constexpr size_t SMEM = 2 * QTILE + NKV * 2 * KTILE + (PSMEM ? 2 * PBYTES : 0) + XBYTES + 256 + 1024;
    // TMEM: two buffers; a buffer holds S (KT fp32 columns), later P (KT/2 columns of bf16 pairs) and O~ (D columns)
    static constexpr int COL_O = KT / 2;
    static constexpr int BUF_COLS = (KT / 2 + D <= 64 && KT <= 64) ? 64 : 128;
    static constexpr int TMEM_COLS = 2 * BUF_COLS;
    // registers after setmaxnreg: 128 * (SOFTMAX + CONTROL) = the CTA's share of the register file
    static constexpr int LAUNCH_REGS = (65536 / (CPS * THREADS)) / 8 * 8;
    static constexpr int CONTROL_REGS = (RS_ > 1) ? 32 : (CPS == 2) ? 56 : 24;
    // what the control warpgroup releases is shared by the 4 RS softmax warps
    static constexpr int SOFTMAX_REGS = (LAUNCH_REGS + (LAUNCH_REGS - CONTROL_REGS) / RS_) / 8 * 8;
    static_assert(COL_O + D <= BUF_COLS && KT <= BUF_COLS, "TMEM buffer layout");
    static_assert(CPS * TMEM_COLS <= 512, "TMEM budget");
};

__device__ __forceinline__ float fmax3f(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

struct TcfCursor {
    int item, it, qi, bh, j, n, g;   // it: ordinal of the item in this CTA; n: tiles of the item; g: tiles so far
};

// ABL: timing-only ablations (results wrong): 1 no max pass, 2 no MUFU (multiply instead), 4 no P store, 8 no O~ load
// TRACED: the event timeline is compiled in (a separate instantiation: even a never-taken trace branch per event costs
// the default kernel 3-7 %).
template <int D, bool DROP, bool PSMEM, int KT_, int CPS_, int ABL = 0, int RS = 1, bool TRACED = false>
__global__ void __launch_bounds__((4 * RS + 4) * 32, CPS_)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                   __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int T, int H, int BH, int nq,
                   float scale_log2, AttnDropKey drop, long long* __restrict__ trace) {
    using C = TcfCfg<D, PSMEM, KT_, CPS_, RS>;
    static_assert(RS == 1 || (RS == 2 && !PSMEM && KT_ == 128), "row split: two threads per row of a 128-key tile, P in TMEM");
    constexpr int KT = C::KT, RB = C::RB, QTILE = C::QTILE, KTILE = C::KTILE, NKV = C::NKV, NCH = KT / 32 / RS;
    constexpr int KW = KT / RS;                                 // keys per thread and tile
    constexpr int FOLD_AT = (NCH >= 4) ? 0 : -1;            // fold O~ of the previous tile after this 32-key chunk (-1: after the maximum pass)
    constexpr int SW = 4 * RS;                                  // softmax warps; then: S issuer, TMA producer, TMEM allocation, P V issuer
    constexpr uint32_t LT = umma_layout_for_row_bytes(RB);
    constexpr uint32_t COL_O = C::COL_O, BUFC = C::BUF_COLS;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // everything below is a 32-bit shared-space address (no generic -> shared conversions in the loops)
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = smem;                               // [2][QTILE]
    const uint32_t sK = sQ + 2 * QTILE;                     // [NKV][KTILE]
    const uint32_t sV = sK + NKV * KTILE;                   // [NKV][KTILE]
    const uint32_t sP = sV + NKV * KTILE;                   // [2][PBYTES] (PSMEM only)
    const uint32_t sX = sP + (PSMEM ? 2 * C::PBYTES : 0);   // [3][RS][128] floats (row split only)
    const uint32_t bars = sX + C::XBYTES;
    const uint32_t bar_q_full = bars;                       // [2] Q tile landed
    const uint32_t bar_q_free = bars + 16;                  // [2] the item's last S MMA has read it
    const uint32_t bar_kv_full = bars + 32;                 // [NKV]
    const uint32_t bar_kv_free = bar_kv_full + 8 * NKV;     // [NKV] the tile's P V MMAs have read it
    const uint32_t bar_s_full = bar_kv_free + 8 * NKV;      // [2] S complete in TMEM
    const uint32_t bar_p_full = bar_s_full + 16;            // [2] P written by every softmax thread
    const uint32_t bar_o_full = bar_s_full + 32;            // [2] O~ complete in TMEM
    const uint32_t bar_buf_free = bar_s_full + 48;          // [2] O~ read: the buffer may take S of tile g + 2
    const uint32_t tmem_slot = bar_s_full + 64;

    const int E = H * D;
    const int items = nq * BH;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

    // Diagnostic timeline (cb200_set_attention_trace): lane 0 of every warp of CTA 0 appends (event << 40 | clock) words
    long long* tr = (TRACED && trace != nullptr && blockIdx.x == 0 && lane == 0) ? trace + warp * 512 : nullptr;
    int tr_n = 0;
    auto TR = [&](int ev) {
        if (TRACED && tr != nullptr && tr_n < 511) tr[++tr_n] = (static_cast<long long>(ev) << 40) | (clock64() & 0xFFFFFFFFFFll);
    };
    TR(1);

    if (warp == SW && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_kv);
        for (int i = 0; i < 2; ++i) {
            mbar_init_a(bar_q_full + 8 * i, 1);
            mbar_init_a(bar_q_free + 8 * i, 1);
            mbar_init_a(bar_s_full + 8 * i, 1);
            mbar_init_a(bar_p_full + 8 * i, 128 * RS);
            mbar_init_a(bar_o_full + 8 * i, 1);
            mbar_init_a(bar_buf_free + 8 * i, 128 * RS);
        }
        for (int i = 0; i < NKV; ++i) {
            mbar_init_a(bar_kv_full + 8 * i, 1);
            mbar_init_a(bar_kv_free + 8 * i, 1);
        }
        mbar_fence_init();
    }
    if (warp == SW + 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");

    auto setup = [&](TcfCursor& c) {
        c.qi = nq - 1 - c.item / BH;                        // heaviest query tiles first
        c.bh = c.item % BH;
        c.n = (c.qi + 1) * (128 / KT);
        c.j = 0;
    };
    auto advance = [&](TcfCursor& c) {
        ++c.g;
        if (++c.j == c.n) {
            c.item += gridDim.x;
            ++c.it;
            if (c.item < items) setup(c);
        }
    };

    if (warp >= SW) {
        setmaxnreg_dec<C::CONTROL_REGS>();
        if (warp == SW + 1) {
            // ===================== TMA producer =====================
            if (elect_one()) {
                TcfCursor c{static_cast<int>(blockIdx.x), 0, 0, 0, 0, 0, 0};
                if (c.item < items) setup(c);
                while (c.item < items) {
                    const int b = c.bh / H, h = c.bh % H;
                    const int row0 = b * T;
                    if (c.j == 0) {
                        const int qs = c.it & 1;
                        mbar_wait_a(bar_q_free + 8 * qs, ((c.it >> 1) & 1) ^ 1);
                        mbar_expect_tx_a(bar_q_full + 8 * qs, QTILE);
                        tma_load_2d_a(sQ + qs * QTILE, &tm_q, bar_q_full + 8 * qs, h * D, row0 + c.qi * 128);
                    }
                    const int st = c.g % NKV;
                    mbar_wait_a(bar_kv_free + 8 * st, ((c.g / NKV) & 1) ^ 1);
                    mbar_expect_tx_a(bar_kv_full + 8 * st, 2 * KTILE);
                    tma_load_2d_a(sK + st * KTILE, &tm_kv, bar_kv_full + 8 * st, E + h * D, row0 + c.j * KT);
                    tma_load_2d_a(sV + st * KTILE, &tm_kv, bar_kv_full + 8 * st, 2 * E + h * D, row0 + c.j * KT);
                    advance(c);
                }
            }
        } else if (warp == SW) {
            // ===================== S issuer (runs one tile ahead of the softmax) =====================
            // Two issuing threads (S here, P V in warp 7): an issuing thread runs alone at one instruction every few
            // cycles, and with 9 small MMAs per tile that instruction stream is a large part of a tile's critical path.
            // Descriptors are (lo, hi) words: hi is constant, lo = start address field + constant offsets.
            if (elect_one()) {
                constexpr uint32_t IDESC_S = umma_idesc_bf16(128, KT, 0, 0);     // Q K^T
                constexpr uint32_t HI_T = umma_desc_hi(8 * RB, LT);
                const uint32_t q_lo = umma_desc_lo(sQ, 16), k_lo = umma_desc_lo(sK, 16);
                TcfCursor c{static_cast<int>(blockIdx.x), 0, 0, 0, 0, 0, 0};
                if (c.item < items) setup(c);
                while (c.item < items) {
                    const int qs = c.it & 1, st = c.g % NKV, buf = c.g & 1;
                    if (c.j == 0) mbar_wait_a(bar_q_full + 8 * qs, (c.it >> 1) & 1);
                    mbar_wait_a(bar_kv_full + 8 * st, (c.g / NKV) & 1);
                    mbar_wait_a(bar_buf_free + 8 * buf, ((c.g >> 1) & 1) ^ 1);
                    tc_fence_after();
                    TR(10);
                    const uint32_t aq = q_lo + qs * (QTILE >> 4), ak = k_lo + st * (KTILE >> 4);
#pragma unroll
                    for (int ks = 0; ks < D / 16; ++ks)
                        umma_bf16_w(tmem + buf * BUFC, aq + ks * 2, HI_T, ak + ks * 2, HI_T, IDESC_S, ks > 0 ? 1u : 0u);
                    umma_commit_a(bar_s_full + 8 * buf);
                    if (c.j == c.n - 1) umma_commit_a(bar_q_free + 8 * qs);
                    TR(11);
                    advance(c);
                }
            }
        } else if (warp == SW + 3) {
            // ===================== P V issuer =====================
            if (elect_one()) {
                constexpr uint32_t IDESC_O = umma_idesc_bf16(128, D, 0, 1);      // P V (V is MN-major)
                constexpr uint32_t HI_T = umma_desc_hi(8 * RB, LT);
                constexpr uint32_t HI_P = umma_desc_hi(1024, 2u);
                const uint32_t v_lo = umma_desc_lo(sV, KT * RB), p_lo = umma_desc_lo(sP, 16);
                TcfCursor c{static_cast<int>(blockIdx.x), 0, 0, 0, 0, 0, 0};
                if (c.item < items) setup(c);
                while (c.item < items) {
                    const int st = c.g % NKV, buf = c.g & 1;
                    mbar_wait_a(bar_kv_full + 8 * st, (c.g / NKV) & 1);          // (long complete: S of the tile used it)
                    mbar_wait_a(bar_p_full + 8 * buf, (c.g >> 1) & 1);
                    tc_fence_after();
                    TR(20);
                    const uint32_t av = v_lo + st * (KTILE >> 4);
#pragma unroll
                    for (int ks = 0; ks < KT / 16; ++ks) {
                        if (PSMEM)
                            umma_bf16_w(tmem + buf * BUFC + COL_O, p_lo + buf * (C::PBYTES >> 4) + (ks >> 2) * (16384 >> 4) + (ks & 3) * 2,
                                        HI_P, av + ks * (16 * RB >> 4), HI_T, IDESC_O, ks > 0 ? 1u : 0u);
                        else
                            umma_bf16_ts_w(tmem + buf * BUFC + COL_O, tmem + buf * BUFC + ks * 8, av + ks * (16 * RB >> 4), HI_T,
                                           IDESC_O, ks > 0 ? 1u : 0u);
                    }
                    umma_commit_a(bar_o_full + 8 * buf);
                    umma_commit_a(bar_kv_free + 8 * st);
                    TR(21);
                    advance(c);
                }
            }
        }
    } else {
        // ===================== softmax warps: thread = query row (RS = 2: half of one) =====================
        setmaxnreg_inc<C::SOFTMAX_REGS>();
        const int quad = warp & 3;                          // TMEM lane quadrant = 32-row band of the tile
        const int part = warp >> 2;                         // which KW-key part of a row this thread owns
        const int r = quad * 32 + lane;
        const uint32_t t_lane = tmem + (static_cast<uint32_t>(quad * 32) << 16);
        const float thr = __uint_as_float(drop.thr_bits);
        const float ks_scale = DROP ? drop.keep_scale : 1.0f;
        const uint64_t c2 = f2_pack(scale_log2, scale_log2);
        constexpr int DW = D / RS;                          // output columns folded and stored by this thread
        // exchange between the RS threads of a row: a named barrier per lane quadrant, values through shared memory
        auto exchange = [&](int slot, float mine) -> float {
            if (RS == 1) return mine;
            const uint32_t base = sX + static_cast<uint32_t>(slot * RS * 128 * 4);
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + (part * 128 + r) * 4), "f"(mine) : "memory");
            asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(32 * RS) : "memory");
            float other;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(other) : "r"(base + ((part ^ 1) * 128 + r) * 4) : "memory");
            return other;
        };
        int g = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int qi = nq - 1 - item / BH, bh = item % BH;
            const int b = bh / H, h = bh % H;
            const int n = (qi + 1) * (128 / KT);
            const int row_g = qi * 128 + r;                 // query row inside the sequence
            const int rmin = qi * 128 + quad * 32, rmax = rmin + 31;
            const uint32_t dbase = DROP ? attn_drop_base(drop, bh) : 0u;
            float o_acc[DW];
#pragma unroll
            for (int d = 0; d < DW; ++d) o_acc[d] = 0.f;
            float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;

            auto fold_o = [&](int gp) {                     // o = o * alpha + O~ of tile gp, then release its buffer
                const int pb = gp & 1;
                mbar_wait_a(bar_o_full + 8 * pb, (gp >> 1) & 1);
                tc_fence_after();
                TR(35);
                uint32_t op[DW];
                if (!(ABL & 8)) {
                    if (DW == 8) tmem_ld8(t_lane + pb * BUFC + COL_O + part * DW, *reinterpret_cast<uint32_t(*)[8]>(&op[0]));
                    else {
#pragma unroll
                        for (int d0 = 0; d0 < DW; d0 += 16)
                            tmem_ld16(t_lane + pb * BUFC + COL_O + part * DW + d0, *reinterpret_cast<uint32_t(*)[16]>(&op[d0 % DW]));
                    }
                }
                tmem_ld_wait();
#pragma unroll
                for (int d = 0; d < DW; ++d) o_acc[d] = fmaf(o_acc[d], alpha_prev, __uint_as_float(op[d]));
                tc_fence_before();
                mbar_arrive_a(bar_buf_free + 8 * pb);
            };

            for (int j = 0; j < n; ++j, ++g) {
                const int buf = g & 1;
                const uint32_t tbuf = t_lane + buf * BUFC;
                mbar_wait_a(bar_s_full + 8 * buf, (g >> 1) & 1);
                tc_fence_after();
                TR(30);
                uint32_t s[KW];
#pragma unroll
                for (int c = 0; c < NCH; ++c) tmem_ld32(tbuf + part * KW + 32 * c, *reinterpret_cast<uint32_t(*)[32]>(&s[32 * c]));
                tmem_ld_wait();
                TR(31);

                const int key0 = j * KT + part * KW;        // first key of this thread's part
                const bool diag = j * KT + KT - 1 > qi * 128;   // some key of the tile lies above some row of the q tile
                uint32_t x = 0;
                if (DROP) {
                    x = attn_row_seed(dbase, static_cast<uint32_t>(row_g), static_cast<uint32_t>(key0 >> 7));
                    if (key0 & 64) x *= mcg_mul_pow(32);
                }
                float alpha = 0.f;
                uint64_t sum_a = f2_pack(0.f, 0.f), sum_b = sum_a;
                // Two copies of the tile arithmetic: only tiles that touch the diagonal pay for the causal mask
                // (32-key chunks wholly above the warp's rows are skipped, the straddling ones are masked per element).
                auto tile_math = [&](auto masked) {
                    constexpr bool MASKED = decltype(masked)::value;
                    // ---- running max ----
                    float mx0 = (RS == 1) ? m_run : -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
                    for (int c = 0; c < ((ABL & 1) ? 0 : NCH); ++c) {
                        const int kmin = key0 + 32 * c;
                        if (MASKED && kmin > rmax) continue;
                        if (MASKED && kmin + 31 > rmin) {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (kmin + i > row_g) s[32 * c + i] = 0xff800000u;   // -inf
                        }
#pragma unroll
                        for (int i = 0; i < 32; i += 8) {
                            mx0 = fmax3f(mx0, __uint_as_float(s[32 * c + i]), __uint_as_float(s[32 * c + i + 1]));
                            mx1 = fmax3f(mx1, __uint_as_float(s[32 * c + i + 2]), __uint_as_float(s[32 * c + i + 3]));
                            mx2 = fmax3f(mx2, __uint_as_float(s[32 * c + i + 4]), __uint_as_float(s[32 * c + i + 5]));
                            mx3 = fmax3f(mx3, __uint_as_float(s[32 * c + i + 6]), __uint_as_float(s[32 * c + i + 7]));
                        }
                    }
                    float m_new = fmaxf(fmax3f(mx0, mx1, mx2), mx3);
                    if (RS > 1) m_new = fmax3f(m_run, m_new, exchange(g & 1, m_new));   // the row's other part
                    if (ABL & 1) m_new = 8.f;
                    alpha = fast_exp2((m_run - m_new) * scale_log2);   // first tile: exp2(-inf) = 0
                    m_run = m_new;
                    TR(32);
                    // O~ of the previous tile is folded here, after the maximum pass: its P V MMAs were triggered at the
                    // end of the previous tile and need ~600 cycles (wake-up of the issuing thread, 8 MMAs, commit);
                    // at the top of the tile the softmax warps waited ~300 of them (tools/trace_attention_fwd.py)
                    if (FOLD_AT < 0 && j > 0) fold_o(g - 1);
                    const float nmc = -m_new * scale_log2;
                    const uint64_t n2 = f2_pack(nmc, nmc);
                    // ---- P = exp2(S c - m c), row sum (before dropout), dropout, bf16 ----
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        const int kmin = key0 + 32 * c;
                        uint32_t pk[16];
                        if (MASKED && kmin > rmax) {
#pragma unroll
                            for (int q = 0; q < 16; ++q) pk[q] = 0u;
                        } else {
#pragma unroll
                            for (int q = 0; q < 16; ++q) {
                                const uint64_t t2 = f2_fma(f2_pack(s[32 * c + 2 * q], s[32 * c + 2 * q + 1]), c2, n2);
                                float t0, t1;
                                f2_unpack(t2, t0, t1);
                                float p0 = (ABL & 2) ? t0 * 0.001f : fast_exp2(t0), p1 = (ABL & 2) ? t1 * 0.001f : fast_exp2(t1);
                                uint64_t p2 = f2_pack(p0, p1);
                                if (q & 1) sum_b = f2_add(sum_b, p2);
                                else       sum_a = f2_add(sum_a, p2);
                                if (DROP) {
                                    float m0, m1;
                                    attn_drop_pair(x, thr, m0, m1);
                                    p2 = f2_mul(p2, f2_pack(m0, m1));
                                    f2_unpack(p2, p0, p1);
                                }
                                pk[q] = pack_bf16(p0, p1);
                            }
                        }
                        if (PSMEM) {
                            // row r of the 64-key half: 128 bytes, 16-byte pieces XOR-swizzled by (r & 7) (SWIZZLE_128B)
                            const uint32_t base = sP + buf * C::PBYTES + (c >> 1) * 16384 + r * 128;
#pragma unroll
                            for (int q4 = 0; q4 < 4; ++q4) {
                                const uint32_t off = static_cast<uint32_t>((((c & 1) * 4 + q4) ^ (r & 7)) << 4);
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + off), "r"(pk[4 * q4]),
                                             "r"(pk[4 * q4 + 1]), "r"(pk[4 * q4 + 2]), "r"(pk[4 * q4 + 3]) : "memory");
                            }
                        } else {
                            if (!(ABL & 4)) tmem_st16(tbuf + part * (KW / 2) + 16 * c, pk);
                            else asm volatile("" ::"r"(pk[0] ^ pk[5] ^ pk[10] ^ pk[15]));
                        }
                        if (FOLD_AT == c && j > 0) fold_o(g - 1);
                    }
                };
                if (diag) tile_math(std::true_type{});
                else      tile_math(std::false_type{});
                float sa0, sa1, sb0, sb1;
                f2_unpack(sum_a, sa0, sa1);
                f2_unpack(sum_b, sb0, sb1);
                l_run = fmaf(l_run, alpha, (sa0 + sa1) + (sb0 + sb1));     // (RS = 2: the sum over this thread's part)
                if (PSMEM) {
                    fence_proxy_async_smem();
                } else {
                    TR(33);
                    tmem_st_wait();
                    tc_fence_before();
                }
                mbar_arrive_a(bar_p_full + 8 * buf);
                TR(34);
                alpha_prev = alpha;
            }
            fold_o(g - 1);
            if (RS > 1) l_run += exchange(2, l_run);

            // ---- finalize the row ----
            if (row_g < T) {
                const float inv = ks_scale / l_run;
                __nv_bfloat16* dst = out + (static_cast<size_t>(b) * T + row_g) * E + h * D + part * DW;
#pragma unroll
                for (int d0 = 0; d0 < DW; d0 += 8) {
                    uint4 v;
                    v.x = pack_bf16(o_acc[d0] * inv, o_acc[d0 + 1] * inv);
                    v.y = pack_bf16(o_acc[d0 + 2] * inv, o_acc[d0 + 3] * inv);
                    v.z = pack_bf16(o_acc[d0 + 4] * inv, o_acc[d0 + 5] * inv);
                    v.w = pack_bf16(o_acc[d0 + 6] * inv, o_acc[d0 + 7] * inv);
                    *reinterpret_cast<uint4*>(dst + d0) = v;
                }
                if (lse != nullptr && part == 0) lse[(static_cast<size_t>(b) * H + h) * T + row_g] = fmaf(m_run, scale_log2, log2f(l_run));
            }
        }
    }

    if (TRACED && tr != nullptr) tr[0] = tr_n;
    tc_fence_before();
    __syncthreads();
    if (warp == SW + 2) {
        tc_fence_after();
        tmem_dealloc<C::TMEM_COLS>(tmem);
    }
}

template <int D, bool DROP, bool PSMEM, int KT, int CPS, int ABL = 0, int RS = 1, bool TRACED = false>
static int launch_fwd_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse, int B, int T, int H, float scale,
                         const AttnDropKey& key, cudaStream_t s) {
    using C = TcfCfg<D, PSMEM, KT, CPS, RS>;
    constexpr size_t smem = C::SMEM;
    const int E = H * D;
    CUtensorMap tm_q, tm_kv;
    int rc = make_tmap_bf16_sw(&tm_q, qkv, 3 * E, static_cast<uint64_t>(B) * T, 3 * E, D, 128, C::RB);
    if (rc) return rc;
    rc = make_tmap_bf16_sw(&tm_kv, qkv, 3 * E, static_cast<uint64_t>(B) * T, 3 * E, D, C::KT, C::RB);
    if (rc) return rc;
    auto kernel = attn_fwd_tc_kernel<D, DROP, PSMEM, KT, CPS, ABL, RS, TRACED>;
    static bool configured = false;
    if (!configured) {
        CB200_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    // CTAs per SM: CPS by registers and TMEM (by construction), fewer when shared memory does not allow it
    int per_sm = CPS;
    while (per_sm > 1 && per_sm * (smem + 1024) > 233472) --per_sm;
    const int nq = (T + 127) / 128;
    const long long items = static_cast<long long>(nq) * B * H;
    long long grid = static_cast<long long>(per_sm) * device_sm_count();
    if (grid > items) grid = items;
    kernel<<<static_cast<int>(grid), C::THREADS, smem, s>>>(tm_q, tm_kv, out, lse, T, H, B * H, nq,
                                                              scale * 1.4426950408889634f, key, attention_get_trace());
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

template <int D, int KT, int CPS>
static int launch_fwd_tc_d(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse, int B, int T, int H, float scale,
                           const AttnDropKey& key, bool psmem, cudaStream_t s) {
    const bool dropping = key.thr_bits != 0;
    if (psmem)
        return dropping ? launch_fwd_tc<D, true, true, KT, 2>(qkv, out, lse, B, T, H, scale, key, s)
                        : launch_fwd_tc<D, false, true, KT, 2>(qkv, out, lse, B, T, H, scale, key, s);
    return dropping ? launch_fwd_tc<D, true, false, KT, CPS>(qkv, out, lse, B, T, H, scale, key, s)
                    : launch_fwd_tc<D, false, false, KT, CPS>(qkv, out, lse, B, T, H, scale, key, s);
}

// The tcgen05 forward.  variant 0: P stays in TMEM (TS-form MMA), the default tile shape; 2: P goes through shared
// memory; 3 / 4 / 5 (d_h 16 only, A/B): P in TMEM with 64-key tiles and 3 / 4 CTAs per SM, 128-key tiles and 2 per SM.
int attention_fwd_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse, int B, int T, int H, int D, float scale,
                     const AttnDropKey& key, int variant, cudaStream_t s) {
    const bool psmem = variant == 2;
    switch (D) {
        case 16:
            if (variant == 3) return launch_fwd_tc_d<16, 64, 3>(qkv, out, lse, B, T, H, scale, key, false, s);
            if (variant == 4) return launch_fwd_tc_d<16, 64, 4>(qkv, out, lse, B, T, H, scale, key, false, s);
            if (variant == 0 && attention_get_trace() != nullptr)      // diagnostic: the default shape with the event timeline
                return key.thr_bits != 0 ? launch_fwd_tc<16, true, false, 128, 2, 0, 1, true>(qkv, out, lse, B, T, H, scale, key, s)
                                         : launch_fwd_tc<16, false, false, 128, 2, 0, 1, true>(qkv, out, lse, B, T, H, scale, key, s);
            if (variant == 7)       // two threads per row: 8 softmax warps per CTA
                return key.thr_bits != 0 ? launch_fwd_tc<16, true, false, 128, 2, 0, 2>(qkv, out, lse, B, T, H, scale, key, s)
                                         : launch_fwd_tc<16, false, false, 128, 2, 0, 2>(qkv, out, lse, B, T, H, scale, key, s);
            if (variant >= 100) {   // timing-only ablations of the default shape, dropout off
                switch (variant - 100) {
                    case 1: return launch_fwd_tc<16, false, false, 128, 2, 1>(qkv, out, lse, B, T, H, scale, key, s);
                    case 2: return launch_fwd_tc<16, false, false, 128, 2, 2>(qkv, out, lse, B, T, H, scale, key, s);
                    case 3: return launch_fwd_tc<16, false, false, 128, 2, 3>(qkv, out, lse, B, T, H, scale, key, s);
                    case 4: return launch_fwd_tc<16, false, false, 128, 2, 4>(qkv, out, lse, B, T, H, scale, key, s);
                    case 8: return launch_fwd_tc<16, false, false, 128, 2, 8>(qkv, out, lse, B, T, H, scale, key, s);
                    case 15: return launch_fwd_tc<16, false, false, 128, 2, 15>(qkv, out, lse, B, T, H, scale, key, s);
                    default: break;
                }
            }
            return launch_fwd_tc_d<16, 128, 2>(qkv, out, lse, B, T, H, scale, key, psmem, s);
        case 32: return launch_fwd_tc_d<32, 64, 2>(qkv, out, lse, B, T, H, scale, key, psmem, s);
        case 64: return launch_fwd_tc_d<64, 64, 2>(qkv, out, lse, B, T, H, scale, key, psmem, s);
        default: break;
    }
    set_error("attention head size %d is not supported (16, 32 or 64)", D);
    return -1;
}

}  // namespace cb200
