// Helpers shared by the tensor-core backward kernels (edge_layer_bwd_tc.cu, virtual_layer_bwd_tc.cu): a thread owns one
// row of a 128-row tile and holds the whole 64-wide row in registers.
#pragma once
#include <cuda_fp16.h>

#include "bwd_common.cuh"
#include "common.cuh"
#include "tc16.cuh"
#include "umma.cuh"

namespace degnn {

__device__ __forceinline__ float row_absmax(const float (&v)[64], float fm = 0.f) {
#pragma unroll
    for (int j = 0; j < 64; ++j) fm = fmaxf(fm, fabsf(v[j]));
    return fm;
}
// power-of-two scale that brings a row maximum `fm` into [2^13, 2^14) (rows of zeros / non-finite maxima keep 1) + inverse
__device__ __forceinline__ void row_scale(float fm, float& s, float& inv) {
    const uint32_t eb = (__float_as_uint(fm) >> 23) & 0xffu;
    const bool live = eb > 0u && eb < 255u;
    const uint32_t sb = live ? min(max(267u - eb, 1u), 254u) : 127u;        // biased exponent of the scale
    s = __uint_as_float(sb << 23);
    inv = __uint_as_float((254u - sb) << 23);
}
// Encode a whole 64-wide row held in registers, times the power-of-two `s`, into the A operand.
__device__ __forceinline__ void encode_row_scaled(const float (&v)[64], float s, uint32_t ta_hi, uint32_t ta_lo) {
    __half2 mx = __floats2half2_rn(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        f32x2 p[8];
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) p[j] = pk2(v[16 * c + 2 * j], v[16 * c + 2 * j + 1]);
        tc16::split16p<true>(p, s, hi, lo, mx);
        umma::tmem_st8(ta_hi + 8 * c, hi);
        umma::tmem_st8(ta_lo + 8 * c, lo);
    }
}
// Encode a whole 64-wide row held in registers into the A operand with its own power-of-two scale; returns 1/scale.
__device__ __forceinline__ float encode_row_regs(const float (&v)[64], uint32_t ta_hi, uint32_t ta_lo) {
    float s, inv;
    row_scale(row_absmax(v), s, inv);
    __half2 mx = __floats2half2_rn(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        f32x2 p[8];
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) p[j] = pk2(v[16 * c + 2 * j], v[16 * c + 2 * j + 1]);
        tc16::split16p<true>(p, s, hi, lo, mx);
        umma::tmem_st8(ta_hi + 8 * c, hi);
        umma::tmem_st8(ta_lo + 8 * c, lo);
    }
    return inv;
}
// one 64-wide fp32 row <-> 64 TMEM columns of the own lane
__device__ __forceinline__ void tmem_store_row(uint32_t taddr, const float (&v)[64]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t d[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) d[j] = __float_as_uint(v[16 * c + j]);
        umma::tmem_st16(taddr + 16 * c, d);
    }
}
__device__ __forceinline__ void tmem_load_row(uint32_t taddr, float (&v)[64]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t d[16];
        umma::tmem_ld16(taddr + 16 * c, d);
        umma::wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[16 * c + j] = __uint_as_float(d[j]);
    }
}
__device__ __forceinline__ void smem_store_row(float* dst, const float (&v)[64]) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
        *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
// Column sums over the 32 rows of a warp of a 64-wide row held in registers (destroys u): after five exchange rounds lane
// l holds the sums of columns 2l and 2l+1 in u[0], u[1] (62 shuffles instead of 64 same-address shared-memory atomics).
__device__ __forceinline__ void warp_colsum64(float (&u)[64], int lane) {
#pragma unroll
    for (int b = 16, n = 64; b >= 1; b >>= 1, n >>= 1) {
        const bool up = lane & b;
        const int half = n >> 1;
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < half) {
                const float lo = u[i], hi = u[i + half];
                const float recv = __shfl_xor_sync(FULL, up ? lo : hi, b);
                u[i] = (up ? hi : lo) + recv;
            }
    }
}
// acc[i][j] += Σ_e Gs[e][n0+i]·Act[e][k0+j], n0 = 8·(t >> 4), k0 = 4·(t & 15): 128 threads cover the 64x64 gradient
__device__ __forceinline__ void wgrad128(float (&acc)[8][4], const float* Gs, const float* Act, int t) {
    const int n0 = 8 * (t >> 4), k0 = 4 * (t & 15);
#pragma unroll 2
    for (int e = 0; e < TILE_M; ++e) {
        const float4 g0 = *reinterpret_cast<const float4*>(Gs + e * LDA + n0);
        const float4 g1 = *reinterpret_cast<const float4*>(Gs + e * LDA + n0 + 4);
        const float4 w = *reinterpret_cast<const float4*>(Act + e * LDA + k0);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            acc[i][0] = fmaf(gg[i], w.x, acc[i][0]);
            acc[i][1] = fmaf(gg[i], w.y, acc[i][1]);
            acc[i][2] = fmaf(gg[i], w.z, acc[i][2]);
            acc[i][3] = fmaf(gg[i], w.w, acc[i][3]);
        }
    }
}
__device__ __forceinline__ void wgrad128_flush(float* g_kmajor, const float (&acc)[8][4], int t) {
    const int n0 = 8 * (t >> 4), k0 = 4 * (t & 15);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(g_kmajor + (k0 + j) * H + n0 + i, acc[i][j]);
}


}  // namespace degnn
