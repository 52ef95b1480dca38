// Micro-benchmark: tensor-pipe cycles per tcgen05.mma (M 128, K 16, bf16) as a function of N and of where / how the
// A operand is read (shared memory K-major, shared memory MN-major, TMEM).  One CTA, one issuing thread; `groups`
// groups of 8 back-to-back MMAs with precomputed descriptors (the issue loop is 3-4 instructions per MMA), one
// commit, wait.  `done cyc` / n is the tensor-pipe time per MMA once it exceeds the issue time.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I composer_b200/csrc
//        tools/cuda/umma_bench.cu -o composer_b200/build/umma_bench
#include <cstdio>
#include "common.cuh"

namespace cb200 { void set_error(const char*, ...) {} void note_launch(long long) {} }
using namespace cb200;

template <int MODE, int N>
__global__ void __launch_bounds__(128, 1) umma_bench_kernel(int groups, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = smem;                 // 32 KB
    const uint32_t sB = smem + 32768;         // 32 KB
    const uint32_t bar = smem + 65536;
    const uint32_t slot = bar + 16;
    const int warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < 16384; i += 128)
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(smem + 4 * i), "r"(0x3c003c00u + (i & 7)) : "memory");
    if (threadIdx.x == 0) { mbar_init_a(bar, 1); mbar_fence_init(); }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot) : "memory");
    if (warp == 0 && elect_one()) {
        constexpr bool BMN = N <= 64;       // B: MN-major tile with rows of 2N bytes, or (N = 128) K-major rows of 32 bytes
        constexpr uint32_t IDESC = umma_idesc_bf16(128, N, MODE == 1 ? 1 : 0, BMN ? 1 : 0);
        constexpr uint32_t HI_B = BMN ? umma_desc_hi(8 * 2 * N, umma_layout_for_row_bytes(2 * N)) : umma_desc_hi(8 * 32, 6u);
        constexpr uint32_t HI_A = MODE == 3 ? umma_desc_hi(8 * 32, 6u) : umma_desc_hi(1024, 2u);
        const uint32_t b_lo = BMN ? umma_desc_lo(sB, 128 * 2 * N) : umma_desc_lo(sB, 16);
        const uint32_t a_lo = MODE == 0 ? umma_desc_lo(sA, 16) : MODE == 1 ? umma_desc_lo(sA, 16384) : umma_desc_lo(sA, 16);
        uint32_t parity = 0;
        for (int rep = 0; rep < reps; ++rep) {
            const long long t0 = clock64();
            for (int g = 0; g < groups; ++g) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint32_t bl = b_lo + (BMN ? ks * (16 * 2 * N >> 4) : 0);
                    const uint32_t acc = (g | ks) ? 1u : 0u;
                    if (MODE == 0) umma_bf16_w(tmem + 256, a_lo + (ks >> 2) * (16384 >> 4) + (ks & 3) * 2, HI_A, bl, HI_B, IDESC, acc);
                    else if (MODE == 1) umma_bf16_w(tmem + 256, a_lo + ks * (2048 >> 4), HI_A, bl, HI_B, IDESC, acc);
                    else if (MODE == 2) umma_bf16_ts_w(tmem + 256, tmem + ks * 8, bl, HI_B, IDESC, acc);
                    else umma_bf16_w(tmem + 256, a_lo, HI_A, bl, HI_B, IDESC, acc);
                }
            }
            umma_commit_a(bar);
            const long long t1 = clock64();
            mbar_wait_a(bar, parity);
            parity ^= 1;
            const long long t2 = clock64();
            if (rep == reps - 1) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

template <int MODE, int N>
static void run(const char* name, long long* out) {
    cudaFuncSetAttribute(umma_bench_kernel<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
    for (int groups : {1, 3, 8, 32}) {
        umma_bench_kernel<MODE, N><<<1, 128, 70000>>>(groups, 3, out);
        long long h[2] = {0, 0};
        cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("%-30s %5d %6d  error %s\n", name, N, groups * 8, cudaGetErrorString(e)); exit(1); }
        printf("%-30s %5d %6d %12lld %12lld %10.1f %10.1f\n", name, N, groups * 8, h[0], h[1], (double)h[0] / (groups * 8),
               (double)h[1] / (groups * 8));
    }
}

int main() {
    long long* out;
    cudaMalloc(&out, 16);
    printf("%-30s %5s %6s %12s %12s %10s %10s\n", "A operand", "N", "n_mma", "issue cyc", "done cyc", "issue/MMA", "done/MMA");
    run<0, 16>("A smem K-major (SW128)", out);  run<0, 64>("A smem K-major (SW128)", out);  run<0, 128>("A smem K-major (SW128)", out);
    run<1, 16>("A smem MN-major (SW128)", out); run<1, 64>("A smem MN-major (SW128)", out);
    run<2, 16>("A in TMEM", out);               run<2, 64>("A in TMEM", out);               run<2, 128>("A in TMEM", out);
    run<3, 16>("A smem K-major 32-byte rows", out); run<3, 64>("A smem K-major 32-byte rows", out); run<3, 128>("A smem K-major 32-byte rows", out);
    return 0;
}
