// Warp-level 3xTF32 tile GEMMs for the backward kernels (mma.sync m16n8k8, fp32 accumulate).
//
// Interim tensor-core path of the backward pass: the fp32-FMA tile GEMMs of the first backward kernels issued
// 2048 FFMA + 768 LDS per thread and GEMM; the same product as 3 x TF32 MMAs (x = hi + lo, hi = top 19 bits,
// a·b ≈ a_lo·b_hi + a_hi·b_lo + a_hi·b_hi — measured 2.4e-6 on O(4) outputs, scripts/umma_selftest.py) needs about
// a quarter of the instructions.  mma.sync is the pre-Blackwell tensor-core interface (it runs on sm_100a, below
// tcgen05 throughput); moving these GEMMs onto tcgen05/TMEM like the forward kernels is the next step.
//
// Tiles live in shared memory as fp32 [128][LDA] (LDA = 68).  8 warps; for the row-wise GEMMs warp w owns rows
// 16w .. 16w+15 and all 64 columns.  Fragment coordinates (lane = 4g + t, g = lane >> 2, t = lane & 3):
//   A 16x8 (row): a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4)        B 8x8 (col): b0 (k=t, n=g)  b1 (k=t+4, n=g)
//   C 16x8:       c0 (g, 2t)  c1 (g, 2t+1)  c2 (g+8, 2t)  c3 (g+8, 2t+1)
// so a thread ends up with rows R0 = 16w+g and R1 = R0+8 and, for n-tile j = 0..7, columns 8j+2t and 8j+2t+1:
//   acc[j][0] = (R0, 8j+2t)  acc[j][1] = (R0, 8j+2t+1)  acc[j][2] = (R1, 8j+2t)  acc[j][3] = (R1, 8j+2t+1).
#pragma once
#include "common.cuh"

namespace degnn {
namespace mma3 {

__device__ __forceinline__ void split(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma3x(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                      const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
    mma(c, al, bh);
    mma(c, ah, bl);
    mma(c, ah, bh);
}
__device__ __forceinline__ void zero(float (&acc)[8][4]) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
}

// acc (+)= As[128 x 64] · W, W[k][n] = Wp[k * pitch + n]  (W in shared OR global memory; GLOBAL selects __ldg)
template <bool GLOBAL>
__device__ __forceinline__ void gemm_rows(float (&acc)[8][4], const float* As, const float* __restrict__ Wp, int pitch,
                                          int warp, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const float* a0p = As + (16 * warp + g) * LDA + t;
#pragma unroll 2
    for (int ks = 0; ks < 8; ++ks) {
        const int k0 = 8 * ks;
        uint32_t ah[4], al[4];
        split(a0p[k0], ah[0], al[0]);
        split(a0p[k0 + 8 * LDA], ah[1], al[1]);
        split(a0p[k0 + 4], ah[2], al[2]);
        split(a0p[k0 + 8 * LDA + 4], ah[3], al[3]);
        const float* w0 = Wp + (k0 + t) * pitch + g;
        const float* w1 = w0 + 4 * pitch;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t bh[2], bl[2];
            split(GLOBAL ? __ldg(w0 + 8 * j) : w0[8 * j], bh[0], bl[0]);
            split(GLOBAL ? __ldg(w1 + 8 * j) : w1[8 * j], bh[1], bl[1]);
            mma3x(acc[j], ah, al, bh, bl);
        }
    }
}

// Weight gradient: D[n][k] += Σ_e Gs[e][n] · Act[e][k] over the 128 rows of the tile (both tiles fp32 [128][LDA]).
// Warp w owns the 16(n) x 32(k) block: n-rows 16·(w & 3) .., k-columns 32·(w >> 2) ..; acc[q] is its q-th 16x8 tile:
//   acc[q][0] = (n = 16(w&3)+g,   k = 32(w>>2)+8q+2t)   acc[q][1] = (same n, k+1)   acc[q][2], acc[q][3]: n + 8.
__device__ __forceinline__ void wgrad(float (&acc)[4][4], const float* Gs, const float* Act, int warp, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const float* gp = Gs + t * LDA + 16 * (warp & 3) + g;
    const float* ap = Act + t * LDA + 32 * (warp >> 2) + g;
#pragma unroll 2
    for (int ks = 0; ks < 16; ++ks) {
        const int e0 = 8 * ks * LDA;
        uint32_t ah[4], al[4];
        split(gp[e0], ah[0], al[0]);
        split(gp[e0 + 8], ah[1], al[1]);
        split(gp[e0 + 4 * LDA], ah[2], al[2]);
        split(gp[e0 + 4 * LDA + 8], ah[3], al[3]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t bh[2], bl[2];
            split(ap[e0 + 8 * q], bh[0], bl[0]);
            split(ap[e0 + 4 * LDA + 8 * q], bh[1], bl[1]);
            mma3x(acc[q], ah, al, bh, bl);
        }
    }
}
// add the warp's block of a weight gradient to the k-major parameter layout: g_kmajor[k * 64 + n]
__device__ __forceinline__ void wgrad_flush(float* g_kmajor, const float (&acc)[4][4], int warp, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const int n = 16 * (warp & 3) + g;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int k = 32 * (warp >> 2) + 8 * q + 2 * t;
        atomicAdd(g_kmajor + k * H + n, acc[q][0]);
        atomicAdd(g_kmajor + (k + 1) * H + n, acc[q][1]);
        atomicAdd(g_kmajor + k * H + n + 8, acc[q][2]);
        atomicAdd(g_kmajor + (k + 1) * H + n + 8, acc[q][3]);
    }
}

}  // namespace mma3
}  // namespace degnn
