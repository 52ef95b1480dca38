// Device helpers shared by the per-step decode kernels (decode.cu) and the
// persistent cluster decode kernel (decode_mega.cu).
#pragma once

#include "common.cuh"

namespace cb200 {

__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ float dot8(const uint4& a, const float (&q)[8]) {
    const float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
    return a0.x * q[0] + a0.y * q[1] + a1.x * q[2] + a1.y * q[3] + a2.x * q[4] + a2.y * q[5] + a3.x * q[6] + a3.y * q[7];
}

// Temperature scale + softmax + multinomial draw for one sequence, executed by one warp on a row of
// logits `z` (global or shared memory).  u ~ U[0,1) from Philox4x32-10 keyed by (seed, global sequence
// index, step), so the tokens do not depend on how sequences are sharded over GPUs.  The draw is the
// inverse CDF in vocabulary order.  greedy selects argmax (first maximum).
__device__ __forceinline__ int sample_row(const float* z, int V, float inv_temperature, int greedy, uint32_t seed_lo,
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
        const float kLog2e = 1.4426950408889634f;
        const float c = inv_temperature * kLog2e;
        // lane owns the contiguous slice [lo, hi) so that the CDF is in vocabulary order
        const int per = (V + 31) / 32;
        const int lo = lane * per, hi = min(V, lo + per);
        float mass = 0.f;
        for (int i = lo; i < hi; ++i) mass += exp2f((z[i] - vmax) * c);
        float prefix = mass;   // inclusive scan over lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float t = __shfl_up_sync(0xffffffffu, prefix, o);
            if (lane >= o) prefix += t;
        }
        const float total = __shfl_sync(0xffffffffu, prefix, 31);
        const Philox4 r = philox4x32_10(step, seq_index, 0x5A17u, 0u, seed_lo, seed_hi);
        u = (r.x >> 8) * (1.0f / 16777216.0f);          // 24-bit uniform in [0, 1)
        const float target = u * total;
        // the lane whose slice [prefix - mass, prefix) contains the target resolves the id
        const float before = prefix - mass;
        const bool mine = (target >= before && target < prefix) || (lane == 31 && target >= prefix);
        int pick = -1;
        if (mine) {
            float run = before;
            pick = max(hi - 1, lo);
            for (int i = lo; i < hi; ++i) {
                run += exp2f((z[i] - vmax) * c);
                if (target < run) { pick = i; break; }
            }
            if (pick >= V) pick = V - 1;
        }
        // lowest lane that claims wins (slices are disjoint; this only breaks float ties)
        const uint32_t ballot = __ballot_sync(0xffffffffu, pick >= 0);
        const int src = ballot ? (__ffs(ballot) - 1) : 0;
        chosen = __shfl_sync(0xffffffffu, pick, src);
        if (chosen < 0) chosen = amax;
    }
    *u_used = u;
    return chosen;
}

}  // namespace cb200
