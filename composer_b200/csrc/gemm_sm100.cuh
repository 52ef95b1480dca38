// Warp-specialised persistent GEMM for sm_100a: TMA -> swizzled smem ->
// tcgen05.mma (accumulators in TMEM) -> tcgen05.ld epilogue.
//
//   D[M, N] (+)= A[M, K] * B[N, K]^T          (logical shapes; "K" = reduced dim)
//
// Either operand may be stored K-major (reduced dimension contiguous, the
// usual "row-major A / col-major B" case) or MN-major (the reduced dimension
// is the slow one: used by the weight-gradient GEMMs, which contract over the
// token dimension of two row-major activation matrices without transposing
// them in HBM).  Tiles are 128 (M) x BN (N) x 64 (K), SWIZZLE_128B.
//
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
// warp 2 = TMEM allocator, warps 4-11 = epilogue: warp w reads TMEM lane quadrant
// w % 4 (thread = accumulator row) and takes the 32-column half (w - 4) / 4 of
// every 64-column slab (the cross-entropy epilogue needs whole rows per thread
// and uses warps 4-7 only).
//
// Reference call sites replaced (composer/models/transformer.py): Conv1D.call
// :194-209 (c_attn :416, attn c_proj :443, c_fc/c_proj :504-505), the tied
// logits matmul :139-144/:818, and their tape.gradient (:920) counterparts.
#pragma once

#include "common.cuh"

namespace cb200 {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 384;   // 4 control warps + 8 epilogue warps

enum GemmEpilogue : int {
    EPI_BIAS_BF16 = 0,      // out0 = bf16(acc + bias)
    EPI_BIAS_GELU = 1,      // out0 = bf16(acc + bias) ; out1 = bf16(gelu(acc + bias))
    EPI_BIAS_DROP_RES = 2,  // out0 = bf16(res + dropout(acc + bias))
    EPI_MUL_DGELU = 3,      // out0 = bf16(acc * gelu'(aux))
    EPI_ATOMIC_F32 = 4,     // outf[m, n] += acc     (split-K weight gradients)
    EPI_CE = 5,             // fused softmax cross-entropy on the accumulator row
};

struct GemmArgs {
    int M, N, K;                  // logical problem
    int num_m_tiles, num_n_tiles; // ceil(M / 128), ceil(N / BN)
    int k_splits;                 // EPI_ATOMIC_F32 only, else 1
    int k_blocks_total;           // ceil(K / 64)
    const float* bias;            // [N] fp32 or nullptr
    const __nv_bfloat16* aux;     // residual (EPI_BIAS_DROP_RES) / pre-activation (EPI_MUL_DGELU), row-major
    int ld_aux;
    float* outf;                  // EPI_ATOMIC_F32 target, row-major; EPI_CE: optional fp32 logits [M, N]
    int ld_outf;
    int valid_m;                  // rows < valid_m are written by EPI_ATOMIC_F32 (others are padding)
    // dropout (EPI_BIAS_DROP_RES)
    DropoutParams drop;
    uint32_t drop_site, drop_layer;
    // cross-entropy (EPI_CE)
    const int32_t* labels;        // [M]
    __nv_bfloat16* dlogits;       // [M, ld_dlogits] or nullptr
    int ld_dlogits;
    float grad_scale;             // dlogits = (softmax - onehot) * grad_scale
    float* loss_sum;              // += sum over rows of (lse - z[y])
    int* correct;                 // += number of rows with argmax == y
};

template <int BN>
struct GemmSmem {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;               // 16 KB
    static constexpr int B_BYTES = BN * GEMM_BK * 2;                    // BN * 128
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGING_BYTES = GEMM_BM * 128;                 // one 64-column bf16 slab
};

// Shared-space accesses with 32-bit addresses: the staging pointers come from an integer-aligned base, so a plain
// dereference compiles to generic LD / ST with 64-bit address arithmetic.
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside a
// SWIZZLE_128B tile whose rows are 128 bytes (what TMA expects to store from).
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk) {
    return row * 128u + ((chunk ^ (row & 7u)) << 4);
}

// KRES > 0: the B operand of this CTA's column tile (KRES k-blocks = the whole K) is loaded once and stays in shared
// memory; the stage ring then carries A only.  With K = 256 a 128 x 256 tile otherwise reads 64 KB of A and 128 KB of
// B (302 MB of L2 -> SM traffic for the 134 MB c_attn GEMM).  Measured: c_attn forward 42.9 -> 38.5 us, the others
// within noise: at K = 256 the epilogue (4 slabs x [TMEM load, bias, pack, swizzled store, 2 CTA barriers, TMA store])
// sets the tile time, not the loads.  The launcher makes the grid a multiple of the number of column tiles, so that
// tile % num_n_tiles, the column tile, is the same for every tile of a CTA.
template <int BN, bool A_MN, bool B_MN, int EPI, int STAGES, int KRES = 0>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_sm100_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmC0, const __grid_constant__ CUtensorMap tmC1,
                  const GemmArgs args) {
    using L = GemmSmem<BN>;
    constexpr int ACC_STAGES = (BN <= 256) ? 2 : 1;
    constexpr int ACC_COLS = (BN <= 256) ? 256 : 512;   // TMEM columns reserved per accumulator stage
    constexpr int TMEM_COLS = 512;
    constexpr int N_CHUNK0 = (BN <= 256) ? BN : 256;    // one tcgen05.mma handles N <= 256
    constexpr int N_CHUNK1 = BN - N_CHUNK0;
    constexpr int B_BOX_ROWS = (BN <= 256) ? BN : BN / 2;   // K-major B: rows per TMA box (<= 256)
    constexpr bool USES_STAGING = (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU || EPI == EPI_BIAS_DROP_RES ||
                                   EPI == EPI_MUL_DGELU);
    constexpr int NUM_OUT = (EPI == EPI_BIAS_GELU) ? 2 : 1;
    // The epilogues that read a second [M, N] operand (residual / pre-activation) fetch its 128 x 64 slab by TMA into
    // the staging buffer the result is then written to (same swizzle, same thread, same 16 bytes), one slab ahead of
    // its use; with thread = row, direct loads would touch 32 different lines per instruction.  tmC1 is its map.
    constexpr bool AUX_TMA = (EPI == EPI_BIAS_DROP_RES || EPI == EPI_MUL_DGELU);
    constexpr int STAGING_BUFS = AUX_TMA ? 3 : 2;
    constexpr int EPI_WARPS = (EPI == EPI_CE) ? 4 : 8;
    static_assert(BN % 16 == 0 && BN <= 512, "bad BN");
    static_assert(!(B_MN && BN > 256), "MN-major B with BN > 256 not supported");

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by SWIZZLE_128B.
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    static_assert(KRES == 0 || (!A_MN && !B_MN && EPI != EPI_ATOMIC_F32), "resident B: K-major operands, no K split");
    constexpr int STAGE_BYTES = KRES ? L::A_BYTES : L::STAGE_BYTES;
    uint8_t* bres = smem;                                  // [KRES][B_BYTES] resident B k-blocks
    uint8_t* stage_base = smem + KRES * L::B_BYTES;
    uint8_t* staging = stage_base + STAGES * STAGE_BYTES; // [STAGING_BUFS][NUM_OUT][16 KB] when USES_STAGING
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + (USES_STAGING ? STAGING_BUFS * NUM_OUT * L::STAGING_BYTES : 0));
    uint64_t* full_bar = bars;                       // [STAGES]
    uint64_t* empty_bar = bars + STAGES;             // [STAGES]
    uint64_t* tmem_full = bars + 2 * STAGES;         // [ACC_STAGES]
    uint64_t* tmem_empty = tmem_full + ACC_STAGES;   // [ACC_STAGES]
    uint64_t* aux_full = tmem_empty + ACC_STAGES;    // [3] aux slab landed (AUX_TMA)
    uint64_t* bres_full = aux_full + 3;              // resident B landed (KRES)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_full + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (USES_STAGING) {
            tma_prefetch_desc(&tmC0);
            if (NUM_OUT == 2 || AUX_TMA) tma_prefetch_desc(&tmC1);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < ACC_STAGES; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], EPI_WARPS * 32);
        }
        for (int s = 0; s < 3; ++s) mbar_init(&aux_full[s], 1);
        mbar_init(bres_full, 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_mn = args.num_m_tiles * args.num_n_tiles;
    const int total_tiles = tiles_mn * args.k_splits;
    const int kb_per_split = (args.k_blocks_total + args.k_splits - 1) / args.k_splits;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            if (KRES > 0 && static_cast<int>(blockIdx.x) < total_tiles) {
                const int n0 = (static_cast<int>(blockIdx.x) % args.num_n_tiles) * BN;
                mbar_expect_tx(bres_full, KRES * L::B_BYTES);
                for (int kb = 0; kb < KRES; ++kb)
#pragma unroll
                    for (int b = 0; b < BN / B_BOX_ROWS; ++b)
                        tma_load_2d(bres + kb * L::B_BYTES + b * B_BOX_ROWS * 128, &tmB, bres_full, kb * GEMM_BK, n0 + b * B_BOX_ROWS);
            }
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int split = tile / tiles_mn;
                const int mn = tile - split * tiles_mn;
                const int m0 = (mn / args.num_n_tiles) * GEMM_BM;
                const int n0 = (mn % args.num_n_tiles) * BN;
                const int kb0 = split * kb_per_split;
                const int kb1 = min(args.k_blocks_total, kb0 + kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = stage_base + stage * STAGE_BYTES;
                    uint8_t* sb = sa + L::A_BYTES;
                    mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                    const int k0 = kb * GEMM_BK;
                    if (A_MN) {
#pragma unroll
                        for (int b = 0; b < GEMM_BM / 64; ++b)
                            tma_load_2d(sa + b * 8192, &tmA, &full_bar[stage], m0 + b * 64, k0);
                    } else {
                        tma_load_2d(sa, &tmA, &full_bar[stage], k0, m0);
                    }
                    if (KRES > 0) {
                        // B is resident
                    } else if (B_MN) {
#pragma unroll
                        for (int b = 0; b < BN / 64; ++b)
                            tma_load_2d(sb + b * 8192, &tmB, &full_bar[stage], n0 + b * 64, k0);
                    } else {
#pragma unroll
                        for (int b = 0; b < BN / B_BOX_ROWS; ++b)
                            tma_load_2d(sb + b * B_BOX_ROWS * 128, &tmB, &full_bar[stage], k0, n0 + b * B_BOX_ROWS);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            constexpr uint32_t IDESC0 = umma_idesc_bf16(GEMM_BM, N_CHUNK0, A_MN ? 1 : 0, B_MN ? 1 : 0);
            constexpr uint32_t IDESC1 = umma_idesc_bf16(GEMM_BM, N_CHUNK1 > 0 ? N_CHUNK1 : 16, A_MN ? 1 : 0, B_MN ? 1 : 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            if (KRES > 0 && static_cast<int>(blockIdx.x) < total_tiles) mbar_wait(bres_full, 0);
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int split = tile / tiles_mn;
                const int kb0 = split * kb_per_split;
                const int kb1 = min(args.k_blocks_total, kb0 + kb_per_split);
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(stage_base + stage * STAGE_BYTES);
                    const uint32_t sb = KRES > 0 ? smem_u32(bres + kb * L::B_BYTES) : sa + L::A_BYTES;
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        // K-major: 16 elements of K = 32 bytes inside the 128-byte swizzle row.
                        // MN-major: 16 K-rows = two 8-row groups of 1024 bytes.
                        const uint64_t da = A_MN ? umma_smem_desc_sw128(sa + k * 2048, 8192, 1024)
                                                 : umma_smem_desc_sw128(sa + k * 32, 16, 1024);
                        const uint64_t db = B_MN ? umma_smem_desc_sw128(sb + k * 2048, 8192, 1024)
                                                 : umma_smem_desc_sw128(sb + k * 32, 16, 1024);
                        const uint32_t accum = (kb > kb0 || k > 0) ? 1u : 0u;
                        umma_bf16(d_tmem, da, db, IDESC0, accum);
                        if (N_CHUNK1 > 0) {
                            // second N chunk: B rows [256, BN) of the K-major tile
                            const uint64_t db1 = umma_smem_desc_sw128(sb + N_CHUNK0 * 128 + k * 32, 16, 1024);
                            umma_bf16(d_tmem + N_CHUNK0, da, db1, IDESC1, accum);
                        }
                    }
                    umma_commit(&empty_bar[stage]);      // frees the smem slot once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[acc]);            // accumulator complete
                if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4 && warp < 4 + EPI_WARPS) {
        // ===================== Epilogue =====================
        const int quad = warp & 3;                   // TMEM lane quadrant this warp may read
        const int half = (warp - 4) >> 2;            // which 32 columns of each 64-column slab
        const int row_in_tile = quad * 32 + lane;
        const int epi_tid = threadIdx.x - 128;
        int acc = 0;
        uint32_t acc_phase = 0;
        int store_parity = 0;                        // which staging buffer pair to use next
        int aux_seq = 0;                             // AUX_TMA: index of the next slab in this CTA's slab sequence
        // AUX_TMA: requests the aux slab of (tile_, slab_) into staging buffer seq % 3 (one thread)
        auto issue_aux = [&](int tile_, int slab_, int seq) {
            const int mn_ = tile_ % tiles_mn;
            const int m0_ = (mn_ / args.num_n_tiles) * GEMM_BM;
            const int n0_ = (mn_ % args.num_n_tiles) * BN + slab_ * 64;
            mbar_expect_tx(&aux_full[seq % 3], L::STAGING_BYTES);
            tma_load_2d(staging + (seq % 3) * L::STAGING_BYTES, &tmC1, &aux_full[seq % 3], n0_, m0_);
        };
        if (AUX_TMA && epi_tid == 0 && static_cast<int>(blockIdx.x) < total_tiles) issue_aux(blockIdx.x, 0, 0);

        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int split = tile / tiles_mn;
            const int mn = tile - split * tiles_mn;
            const int m0 = (mn / args.num_n_tiles) * GEMM_BM;
            const int n0 = (mn % args.num_n_tiles) * BN;
            const int row = m0 + row_in_tile;
            const int kb0 = split * kb_per_split;
            const bool has_work = kb0 < args.k_blocks_total;

            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + acc * ACC_COLS + (static_cast<uint32_t>(quad * 32) << 16);

            if constexpr (USES_STAGING) {
#pragma unroll 1
                for (int slab = 0; slab < BN / 64; ++slab) {
                    const int ncol0 = n0 + slab * 64;
                    if (ncol0 >= args.N) break;      // uniform across the CTA
                    uint8_t* buf0 = staging + ((AUX_TMA ? aux_seq % 3 : store_parity) * NUM_OUT) * L::STAGING_BYTES;
                    uint8_t* buf1 = buf0 + L::STAGING_BYTES;
                    // The accumulator slab and the bias do not depend on the staging buffer: request them before waiting
                    // for it.
                    uint32_t v[32];
                    tmem_ld32(t_row + slab * 64 + half * 32, v);
                    const int c0 = ncol0 + half * 32;
                    float4 bias4[8];
                    if (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU || EPI == EPI_BIAS_DROP_RES) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            bias4[j] = (args.bias != nullptr && c0 + 4 * j < args.N) ? __ldg(reinterpret_cast<const float4*>(args.bias + c0 + 4 * j))
                                                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    if (AUX_TMA) {
                        if (epi_tid == 0) {
                            // next slab of this CTA's sequence: its buffer was last read by the store two slabs ago
                            int nt = tile, ns = slab + 1;
                            if (ns >= BN / 64 || n0 + ns * 64 >= args.N) { nt = tile + gridDim.x; ns = 0; }
                            if (nt < total_tiles) {
                                tma_store_wait_read<1>();
                                issue_aux(nt, ns, aux_seq + 1);
                            }
                        }
                        mbar_wait(&aux_full[aux_seq % 3], (aux_seq / 3) & 1);
                    } else {
                        // the TMA store that last read this buffer pair must have drained
                        if (epi_tid == 0) tma_store_wait_read<NUM_OUT>();
                        epi_bar_sync();
                    }
                    {
                        tmem_ld_wait();
                        float f[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                        if (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU || EPI == EPI_BIAS_DROP_RES) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 b4 = bias4[j >> 2];
                                f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                            }
                        }
                        float g[32];
                        if (EPI == EPI_BIAS_GELU) {
#pragma unroll
                            for (int j = 0; j < 32; j += 2) f2_unpack(gelu_tanh2(f2_pack(f[j], f[j + 1])), g[j], g[j + 1]);
                        }
                        if (EPI == EPI_BIAS_DROP_RES || EPI == EPI_MUL_DGELU) {
                            const bool in_rows = row < args.M;
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                // rows / columns outside the tensor were zero-filled by the TMA load
                                const uint4 a4 = lds128(smem_u32(buf0) + sw128_offset(row_in_tile, half * 4 + (j >> 3)));
                                (void)in_rows;
                                const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w};
                                Philox4 rb{0, 0, 0, 0};
                                if (EPI == EPI_BIAS_DROP_RES && args.drop.threshold16 != 0)
                                    rb = drop_bits_rowmajor(args.drop, args.drop_site, args.drop_layer,
                                                            static_cast<uint32_t>(row), static_cast<uint32_t>((c0 + j) >> 3));
#pragma unroll
                                for (int e = 0; e < 8; e += 2) {
                                    const float2 a2 = unpack_bf16(aw[e >> 1]);
                                    if (EPI == EPI_BIAS_DROP_RES) {
                                        float x0 = f[j + e], x1 = f[j + e + 1];
                                        if (args.drop.threshold16 != 0) {
                                            x0 = (drop_u16(rb, e) < args.drop.threshold16) ? 0.f : x0 * args.drop.keep_scale;
                                            x1 = (drop_u16(rb, e + 1) < args.drop.threshold16) ? 0.f : x1 * args.drop.keep_scale;
                                        }
                                        f[j + e] = a2.x + x0;
                                        f[j + e + 1] = a2.y + x1;
                                    } else {
                                        f2_unpack(f2_mul(f2_pack(f[j + e], f[j + e + 1]), gelu_tanh_grad2(f2_pack(a2.x, a2.y))),
                                                  f[j + e], f[j + e + 1]);
                                    }
                                }
                            }
                        }
                        // swizzled 16-byte stores: conflict-free for thread-per-row
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            const uint32_t chunk = half * 4 + (j >> 3);
                            uint4 o;
                            o.x = pack_bf16(f[j], f[j + 1]); o.y = pack_bf16(f[j + 2], f[j + 3]);
                            o.z = pack_bf16(f[j + 4], f[j + 5]); o.w = pack_bf16(f[j + 6], f[j + 7]);
                            sts128(smem_u32(buf0) + sw128_offset(row_in_tile, chunk), o);
                            if (EPI == EPI_BIAS_GELU) {
                                uint4 o1;
                                o1.x = pack_bf16(g[j], g[j + 1]); o1.y = pack_bf16(g[j + 2], g[j + 3]);
                                o1.z = pack_bf16(g[j + 4], g[j + 5]); o1.w = pack_bf16(g[j + 6], g[j + 7]);
                                sts128(smem_u32(buf1) + sw128_offset(row_in_tile, chunk), o1);
                            }
                        }
                    }
                    fence_proxy_async_smem();
                    epi_bar_sync();
                    if (epi_tid == 0) {
                        tma_store_2d(&tmC0, buf0, ncol0, m0);
                        tma_store_commit();
                        if (NUM_OUT == 2) {
                            tma_store_2d(&tmC1, buf1, ncol0, m0);
                            tma_store_commit();
                        }
                    }
                    store_parity ^= 1;
                    ++aux_seq;
                }
            } else if constexpr (EPI == EPI_ATOMIC_F32) {
                if (has_work) {
#pragma unroll 1
                    for (int c = half * 32; c < BN; c += 64) {
                        if (n0 + c >= args.N) break;
                        uint32_t v[32];
                        tmem_ld32(t_row + c, v);
                        tmem_ld_wait();
                        if (row < args.valid_m) {
                            float* dst = args.outf + static_cast<size_t>(row) * args.ld_outf + n0 + c;
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                if (n0 + c + j < args.N) {
                                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j),
                                                 "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])),
                                                 "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                                                 : "memory");
                                }
                            }
                        }
                    }
                }
            } else {  // EPI_CE : the whole vocabulary row lives in this thread's TMEM lane
                const bool in_rows = row < args.M;
                const int label = (in_rows && args.labels != nullptr) ? args.labels[row] : -1;
                const int V = args.N;
                float vmax = -INFINITY, zy = 0.f;
                int amax = 0;
#pragma unroll 1
                for (int c = 0; c < BN; c += 16) {
                    uint32_t v[16];
                    tmem_ld16(t_row + c, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float z = __uint_as_float(v[j]);
                        const int col = c + j;
                        if (col < V) {
                            if (z > vmax) { vmax = z; amax = col; }   // first maximum wins (tf.argmax)
                            if (col == label) zy = z;
                        }
                    }
                }
                float sum = 0.f;
                const float kLog2e = 1.4426950408889634f;
#pragma unroll 1
                for (int c = 0; c < BN; c += 16) {
                    uint32_t v[16];
                    tmem_ld16(t_row + c, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c + j < V) sum += exp2f((__uint_as_float(v[j]) - vmax) * kLog2e);
                }
                const float inv_sum = 1.0f / sum;
                float row_loss = in_rows ? (logf(sum) + vmax - zy) : 0.f;
                int row_hit = (in_rows && amax == label) ? 1 : 0;
                if (args.dlogits != nullptr || args.outf != nullptr) {
#pragma unroll 1
                    for (int c = 0; c < BN; c += 16) {
                        uint32_t v[16];
                        tmem_ld16(t_row + c, v);
                        tmem_ld_wait();
                        if (in_rows) {
                            if (args.dlogits != nullptr && c < args.ld_dlogits) {
                                float d[16];
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    const int col = c + j;
                                    float p = (col < V) ? exp2f((__uint_as_float(v[j]) - vmax) * kLog2e) * inv_sum : 0.f;
                                    if (col == label) p -= 1.0f;
                                    d[j] = p * args.grad_scale;
                                }
                                uint4 o0, o1;
                                o0.x = pack_bf16(d[0], d[1]); o0.y = pack_bf16(d[2], d[3]);
                                o0.z = pack_bf16(d[4], d[5]); o0.w = pack_bf16(d[6], d[7]);
                                o1.x = pack_bf16(d[8], d[9]); o1.y = pack_bf16(d[10], d[11]);
                                o1.z = pack_bf16(d[12], d[13]); o1.w = pack_bf16(d[14], d[15]);
                                uint4* dst = reinterpret_cast<uint4*>(args.dlogits + static_cast<size_t>(row) * args.ld_dlogits + c);
                                dst[0] = o0;
                                dst[1] = o1;
                            }
                            if (args.outf != nullptr) {
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    if (c + j < V) args.outf[static_cast<size_t>(row) * args.ld_outf + c + j] = __uint_as_float(v[j]);
                            }
                        }
                    }
                }
                row_loss = warp_sum(row_loss);
                row_hit = __reduce_add_sync(0xffffffffu, row_hit);
                if (lane == 0) {
                    if (args.loss_sum != nullptr) atomicAdd(args.loss_sum, row_loss);
                    if (args.correct != nullptr) atomicAdd(args.correct, row_hit);
                }
            }

            // all of this thread's TMEM reads for the tile are complete
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
        if (USES_STAGING && epi_tid == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

}  // namespace cb200
