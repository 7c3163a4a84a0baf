// Shared device code for the DistEGNN sm_100a kernels: the 128x64x64 fp32 tile GEMM every fused
// stage is built from, SiLU, parameter-block offsets, error plumbing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/distegnn_b200.h"

namespace degnn {

constexpr int H = 64;          // hidden width
constexpr int TILE_M = 128;    // rows (edges / node·channel pairs / nodes) per CTA tile
constexpr int LDA = 68;        // smem leading dimension of an activation tile (floats); 68 = 64+4 keeps
                               // float4 alignment and shifts consecutive rows by 4 banks
constexpr int NTHREADS = 256;  // 16 (ty) x 16 (tx); thread owns rows ty+16*i (i<8), cols 4*tx..4*tx+3
constexpr unsigned FULL = 0xffffffffu;

// ---- host side: error string + parameter layout ------------------------------------------------
void set_error(const char* fmt, ...);
struct Layout {
    int64_t off[DISTEGNN_P_NUM_FIELDS];
    int64_t total;
};
Layout make_layout(int A, int C, int Na);
int check_dims(int A, int C, int Na);

#define DEGNN_CHECK_ARG(cond, msg)                                  \
    do {                                                            \
        if (!(cond)) {                                              \
            ::degnn::set_error("%s: %s", __func__, msg);            \
            return DISTEGNN_EINVAL;                                 \
        }                                                           \
    } while (0)

#define DEGNN_CHECK_LAUNCH()                                                             \
    do {                                                                                 \
        cudaError_t e__ = cudaPeekAtLastError();                                         \
        if (e__ != cudaSuccess) {                                                        \
            ::degnn::set_error("%s: CUDA error: %s", __func__, cudaGetErrorString(e__)); \
            return DISTEGNN_ECUDA;                                                       \
        }                                                                                \
    } while (0)

int sm_count();   // cached multiprocessor count of the current device
// Opt `kernel` in to `bytes` of dynamic shared memory, once per (device, kernel).  Not a stream operation, so it is
// done on the first (eager) call and never again — nothing but launches happens under CUDA-graph capture.
void ensure_dynamic_smem(const void* kernel, int bytes);

// ---- device helpers -----------------------------------------------------------------------------
__device__ __forceinline__ float silu(float x) {
#ifdef DEGNN_DIAG_NO_SILU      // diagnostic build only (timing apportionment): NOT the model's activation
    return x * 0.5f;
#endif
    // x·σ(x) = x · rcp(1 + 2^{-x·log2 e}): FMUL, MUFU.EX2, FADD, MUFU.RCP, FMUL — no range fix-ups needed:
    // x → −∞ gives 2^{+big} = inf, rcp(inf) = 0, x·0 = −0;  x → +∞ gives rcp(1) = 1.  |rel err| ≲ 5e-7.
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return x * r;
}
__device__ __forceinline__ float4 silu4(float4 v) {
    return make_float4(silu(v.x), silu(v.y), silu(v.z), silu(v.w));
}
__device__ __forceinline__ float4 ldg4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ float4 fma4(float s, float4 w, float4 a) {
    return make_float4(fmaf(s, w.x, a.x), fmaf(s, w.y, a.y), fmaf(s, w.z, a.z), fmaf(s, w.w, a.w));
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// Copy a k-major [64][64] weight chunk global -> smem (16 KB, 4 float4 per thread).
__device__ __forceinline__ void load_w64(float* Ws, const float* __restrict__ Wg, int tid) {
    const float4* g = reinterpret_cast<const float4*>(Wg);
    float4* s = reinterpret_cast<float4*>(Ws);
#pragma unroll
    for (int i = 0; i < 4; ++i) s[tid + NTHREADS * i] = __ldg(g + tid + NTHREADS * i);
}

// Copy a [rows<=128][64] fp32 row block (row stride 64) global -> As (row stride LDA), optionally
// scaling row r by scale[r] (smem array) — used for agg_m / max(deg,1).  Rows >= rows_valid zeroed.
__device__ __forceinline__ void load_a_tile(float* As, const float* __restrict__ src, int rows_valid,
                                            const float* row_scale, int tid) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int idx = tid + NTHREADS * i;   // 0..2047 float4 slots: row = idx/16, quad = idx%16
        int r = idx >> 4, q = idx & 15;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows_valid) {
            v = ldg4(src + (size_t)r * H + 4 * q);
            if (row_scale) {
                float s = row_scale[r];
                v.x *= s; v.y *= s; v.z *= s; v.w *= s;
            }
        }
        *reinterpret_cast<float4*>(As + r * LDA + 4 * q) = v;
    }
}

__device__ __forceinline__ void zero_acc(float (&acc)[8][4]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
}

// acc[i][j] += Σ_k As[ty+16i][k] · Ws[k][4tx+j]   (128x64 tile, K=64, fp32 FMA)
__device__ __forceinline__ void gemm_tile(float (&acc)[8][4], const float* As, const float* Ws, int ty,
                                          int tx) {
#pragma unroll 2
    for (int k4 = 0; k4 < 16; ++k4) {
        float4 w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            w[j] = *reinterpret_cast<const float4*>(Ws + (4 * k4 + j) * H + 4 * tx);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float4 a = *reinterpret_cast<const float4*>(As + (ty + 16 * i) * LDA + 4 * k4);
            acc[i][0] = fmaf(a.x, w[0].x, acc[i][0]);
            acc[i][1] = fmaf(a.x, w[0].y, acc[i][1]);
            acc[i][2] = fmaf(a.x, w[0].z, acc[i][2]);
            acc[i][3] = fmaf(a.x, w[0].w, acc[i][3]);
            acc[i][0] = fmaf(a.y, w[1].x, acc[i][0]);
            acc[i][1] = fmaf(a.y, w[1].y, acc[i][1]);
            acc[i][2] = fmaf(a.y, w[1].z, acc[i][2]);
            acc[i][3] = fmaf(a.y, w[1].w, acc[i][3]);
            acc[i][0] = fmaf(a.z, w[2].x, acc[i][0]);
            acc[i][1] = fmaf(a.z, w[2].y, acc[i][1]);
            acc[i][2] = fmaf(a.z, w[2].z, acc[i][2]);
            acc[i][3] = fmaf(a.z, w[2].w, acc[i][3]);
            acc[i][0] = fmaf(a.w, w[3].x, acc[i][0]);
            acc[i][1] = fmaf(a.w, w[3].y, acc[i][1]);
            acc[i][2] = fmaf(a.w, w[3].z, acc[i][2]);
            acc[i][3] = fmaf(a.w, w[3].w, acc[i][3]);
        }
    }
}

// act = SiLU(acc + bias) written back to the activation tile (the A operand of the next GEMM).
// Caller must __syncthreads() before (all reads of As done) and after.
__device__ __forceinline__ void bias_silu_to_tile(const float (&acc)[8][4], float4 b, float* As, int ty,
                                                  int tx) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float4 v = make_float4(silu(acc[i][0] + b.x), silu(acc[i][1] + b.y), silu(acc[i][2] + b.z),
                               silu(acc[i][3] + b.w));
        *reinterpret_cast<float4*>(As + (ty + 16 * i) * LDA + 4 * tx) = v;
    }
}

// 1-wide head: out[row] = Σ_col w3[col]·SiLU(acc[row][col] + b[col]); reduced over the 16 tx lanes,
// lane tx==0 stores to dst[row] (smem).
__device__ __forceinline__ void head_dot_to_smem(const float (&acc)[8][4], float4 b, float4 w3,
                                                 float* dst, int ty, int tx) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float v = silu(acc[i][0] + b.x) * w3.x;
        v = fmaf(silu(acc[i][1] + b.y), w3.y, v);
        v = fmaf(silu(acc[i][2] + b.z), w3.z, v);
        v = fmaf(silu(acc[i][3] + b.w), w3.w, v);
        v += __shfl_xor_sync(FULL, v, 1);
        v += __shfl_xor_sync(FULL, v, 2);
        v += __shfl_xor_sync(FULL, v, 4);
        v += __shfl_xor_sync(FULL, v, 8);
        if (tx == 0) dst[ty + 16 * i] = v;
    }
}

}  // namespace degnn
