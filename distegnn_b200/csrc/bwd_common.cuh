// Device helpers shared by the backward kernels (fp32 FMA tile GEMMs on the CUDA cores).
#pragma once
#include "common.cuh"

namespace degnn {

__device__ __forceinline__ float sigmoid_f(float z) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}
// SiLU'(z) = σ(z)·(1 + z·(1 − σ(z)))
__device__ __forceinline__ float dsilu(float z) {
    const float s = sigmoid_f(z);
    return s * fmaf(z, 1.0f - s, 1.0f);
}
// transposed copy of a k-major [64][64] matrix: Wt[n][k] = W[k][n]
__device__ __forceinline__ void load_w64_t(float* Wt, const float* __restrict__ Wg, int tid) {
    for (int i = tid; i < H * H; i += NTHREADS) {
        const int k = i >> 6, n = i & 63;
        Wt[n * H + k] = __ldg(Wg + i);
    }
}
// acc[i][j] += Σ_e Gs[e][n0+i]·Act[e][k0+j] over the 128 rows of the tile   (weight gradient, K = rows)
__device__ __forceinline__ void wgrad_tile(float (&acc)[4][4], const float* Gs, const float* Act, int tid) {
    const int n0 = 4 * (tid >> 4), k0 = 4 * (tid & 15);
#pragma unroll 4
    for (int e = 0; e < TILE_M; ++e) {
        const float4 g = *reinterpret_cast<const float4*>(Gs + e * LDA + n0);
        const float4 w = *reinterpret_cast<const float4*>(Act + e * LDA + k0);
        acc[0][0] = fmaf(g.x, w.x, acc[0][0]); acc[0][1] = fmaf(g.x, w.y, acc[0][1]);
        acc[0][2] = fmaf(g.x, w.z, acc[0][2]); acc[0][3] = fmaf(g.x, w.w, acc[0][3]);
        acc[1][0] = fmaf(g.y, w.x, acc[1][0]); acc[1][1] = fmaf(g.y, w.y, acc[1][1]);
        acc[1][2] = fmaf(g.y, w.z, acc[1][2]); acc[1][3] = fmaf(g.y, w.w, acc[1][3]);
        acc[2][0] = fmaf(g.z, w.x, acc[2][0]); acc[2][1] = fmaf(g.z, w.y, acc[2][1]);
        acc[2][2] = fmaf(g.z, w.z, acc[2][2]); acc[2][3] = fmaf(g.z, w.w, acc[2][3]);
        acc[3][0] = fmaf(g.w, w.x, acc[3][0]); acc[3][1] = fmaf(g.w, w.y, acc[3][1]);
        acc[3][2] = fmaf(g.w, w.z, acc[3][2]); acc[3][3] = fmaf(g.w, w.w, acc[3][3]);
    }
}
// flush a register-resident [n0..n0+3][k0..k0+3] block of a weight gradient into the k-major parameter layout
__device__ __forceinline__ void wgrad_flush(float* g_kmajor, const float (&acc)[4][4], int tid) {
    const int n0 = 4 * (tid >> 4), k0 = 4 * (tid & 15);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(g_kmajor + (k0 + j) * H + n0 + i, acc[i][j]);
}
// elementwise: dst tile = SiLU(src tile)
__device__ __forceinline__ void silu_tile(float* dst, const float* src, int tid) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int idx = tid + NTHREADS * i, r = idx >> 4, q = idx & 15;
        *reinterpret_cast<float4*>(dst + r * LDA + 4 * q) = silu4(*reinterpret_cast<const float4*>(src + r * LDA + 4 * q));
    }
}
__device__ __forceinline__ void red_add_v4(float* gdst, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gdst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}


}  // namespace degnn
