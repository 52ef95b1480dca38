// Pipe micro-benchmark (sm_100a): cycles per warp instruction and SM sub-partition for the instructions the attention
// softmax loops are made of, alone and mixed.  One CTA per SM, W warps per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench pipe_bench.cu && ./pipe_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pack(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float max3(float a, float b, float c) { float d; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ uint32_t cvt2(float a, float b) { uint32_t d; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a)); return d; }

constexpr int N = 64;   // independent chains per thread
constexpr int IT = 256;

template <int MODE>
__global__ void bench(float* out, long long* cycles, float seed) {
    float v[N];
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < IT; ++it) {
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            if (MODE == 0) { v[i] = ex2(v[i]); v[i + 1] = ex2(v[i + 1]); }                       // MUFU only
            if (MODE == 1) { uint64_t p = ffma2(pack(v[i], v[i + 1]), pack(1.0001f, 1.0001f), pack(0.1f, 0.1f)); asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(v[i]), "=f"(v[i + 1]) : "l"(p)); }   // FFMA2 only
            if (MODE == 2) { v[i] = max3(v[i], v[i + 1], seed); }                                // FMNMX3 only
            if (MODE == 3) { acc ^= cvt2(v[i], v[i + 1]); }                                      // F2FP only
            if (MODE == 4) {   // the forward softmax mix per pair: FFMA2, 2 MUFU, FADD2 (as FFMA2), F2FP, 0.5 FMNMX3 x 2
                uint64_t p = ffma2(pack(v[i], v[i + 1]), pack(0.9f, 0.9f), pack(-0.1f, -0.1f));
                float a, b; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p));
                a = ex2(a); b = ex2(b);
                uint64_t s = ffma2(pack(a, b), pack(1.f, 1.f), pack(v[i], v[i + 1]));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(v[i]), "=f"(v[i + 1]) : "l"(s));
                acc ^= cvt2(a, b);
                v[i] = max3(v[i], a, b);
            }
            if (MODE == 5) {   // mix without the MUFU
                uint64_t p = ffma2(pack(v[i], v[i + 1]), pack(0.9f, 0.9f), pack(-0.1f, -0.1f));
                float a, b; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p));
                uint64_t s = ffma2(pack(a, b), pack(1.f, 1.f), pack(v[i], v[i + 1]));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(v[i]), "=f"(v[i + 1]) : "l"(s));
                acc ^= cvt2(a, b);
                v[i] = max3(v[i], a, b);
            }
        }
    }
    const long long t1 = clock64();
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) sum += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum + __uint_as_float(acc);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_instr_count) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&cyc, sizeof(long long));
    for (int warps_per_smsp = 1; warps_per_smsp <= 4; warps_per_smsp *= 2) {
        const int threads = 128 * warps_per_smsp;
        bench<MODE><<<148, threads>>>(out, cyc, 0.5f);
        bench<MODE><<<148, threads>>>(out, cyc, 0.5f);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
        const double n_inst = double(IT) * (N / 2) * per_instr_count * warps_per_smsp;   // warp instructions per SMSP
        printf("%-34s warps/SMSP %d: %8lld cycles, %6.2f cycles per warp instruction (per SMSP)\n", name, warps_per_smsp, c, c / n_inst);
    }
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("MUFU.EX2", 2);
    run<1>("FFMA2", 1);
    run<2>("FMNMX3", 1);
    run<3>("F2FP.BF16 pack", 1);
    run<4>("softmax mix (7 instr / pair)", 7);
    run<5>("mix without MUFU (5 instr / pair)", 5);
    return 0;
}
