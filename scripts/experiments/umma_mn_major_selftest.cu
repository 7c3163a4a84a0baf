// Experiment for the NEXT step of the backward pass: weight-gradient GEMMs D[n][k] = Σ_e G[e][n]·Act[e][k]
// on tcgen05 with BOTH operands MN-major straight from row-major fp16 tiles [e][64] (128-byte rows, 16-byte chunk index
// XORed with e mod 8 = the canonical SWIZZLE_128B MN-major layout), M = 64.
// NOT part of the library (not under distegnn_b200/csrc, not built by build.py).  It has been compile-checked only:
//     nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I distegnn_b200/csrc -I include \
//          scripts/experiments/umma_mn_major_selftest.cu -o /tmp/umma_mn && /tmp/umma_mn
// What it answers when run on a B200:
//   (1) does the descriptor below (layout type 2 = SWIZZLE_128B, SBO = 1024 B per 8 K-rows, start address advanced by
//       2048 B per K = 16 step, a_major = b_major = 1 in the instruction descriptor) reproduce Gᵀ·Act ?
//   (2) where do the 64 rows of an M = 64 accumulator live in TMEM (it dumps all 128 lanes x 64 columns and matches every
//       lane against every reference row).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "umma.cuh"

using namespace degnn::umma;

constexpr int KE = 128, MN = 64;            // K = edges of a tile, M = N = 64

// instruction descriptor kind::f16, fp32 accumulate, A and B MN-major (bits 15, 16), N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t idesc_mn(int M, int N) {
    return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// shared-memory descriptor, SWIZZLE_128B (layout type 2 at [61,64)), version 1 at bit 46
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
        "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}

__global__ void __launch_bounds__(128, 1) selftest(const __half* G, const __half* Act, float* dump) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __half* Gs = reinterpret_cast<__half*>(smem);              // [128 e][64 n], 128-byte rows, chunks swizzled
    __half* As = Gs + KE * MN;
    uint64_t* bar = reinterpret_cast<uint64_t*>(As + KE * MN);
    uint32_t* tbase_s = reinterpret_cast<uint32_t*>(bar + 1);
    const int t = threadIdx.x;
    // thread t owns row e = t of both tiles: 8 chunks of 8 halfs, chunk c stored at position c ^ (e & 7)
    for (int c = 0; c < 8; ++c) {
        const int pos = c ^ (t & 7);
        *reinterpret_cast<uint4*>(Gs + t * MN + 8 * pos) = *reinterpret_cast<const uint4*>(G + t * MN + 8 * c);
        *reinterpret_cast<uint4*>(As + t * MN + 8 * pos) = *reinterpret_cast<const uint4*>(Act + t * MN + 8 * c);
    }
    if (t == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (t < 32) tmem_alloc(tbase_s, 64);
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = *tbase_s;
    if (t == 0) {
        const uint32_t idesc = idesc_mn(64, 64);
        for (int ks = 0; ks < KE / 16; ++ks) {                 // K = 16 per MMA = two 8-row groups = 2048 bytes
            const uint64_t da = desc_sw128(smem_u32(Gs) + ks * 2048, 8192, 1024);
            const uint64_t db = desc_sw128(smem_u32(As) + ks * 2048, 8192, 1024);
            mma_f16_ss(tbase, da, db, idesc, ks > 0);
        }
        mma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    __syncwarp();
    fence_after_sync();
    const uint32_t lane_off = ((uint32_t)(t & ~31)) << 16;     // warp w reads lanes 32w..32w+31
    for (int c = 0; c < 4; ++c) {
        uint32_t d[16];
        tmem_ld16(lane_off + tbase + 16 * c, d);
        wait_ld();
        for (int j = 0; j < 16; ++j) dump[t * 64 + 16 * c + j] = __uint_as_float(d[j]);
    }
    fence_before_sync();
    __syncthreads();
    if (t < 32) tmem_dealloc(tbase, 64);
}

int main() {
    static __half hG[KE * MN], hA[KE * MN];
    static float ref[MN * MN], dump[128 * 64];
    srand(1);
    for (int i = 0; i < KE * MN; ++i) {
        hG[i] = __float2half((rand() % 2001 - 1000) / 1000.0f);
        hA[i] = __float2half((rand() % 2001 - 1000) / 1000.0f);
    }
    for (int n = 0; n < MN; ++n)
        for (int k = 0; k < MN; ++k) {
            double s = 0;
            for (int e = 0; e < KE; ++e) s += (double)__half2float(hG[e * MN + n]) * (double)__half2float(hA[e * MN + k]);
            ref[n * MN + k] = (float)s;
        }
    __half *dG, *dA;
    float* dD;
    cudaMalloc(&dG, sizeof(hG)); cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dD, sizeof(dump));
    cudaMemcpy(dG, hG, sizeof(hG), cudaMemcpyHostToDevice);
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, sizeof(dump));
    const int smem_bytes = 2 * KE * MN * 2 + 64;
    cudaFuncSetAttribute(selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    selftest<<<1, 128, smem_bytes>>>(dG, dA, dD);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
    cudaMemcpy(dump, dD, sizeof(dump), cudaMemcpyDeviceToHost);
    // for every TMEM lane: which reference row (if any) does it hold?
    int found = 0;
    for (int lane = 0; lane < 128; ++lane) {
        int best = -1; double beste = 1e30;
        for (int n = 0; n < MN; ++n) {
            double e = 0;
            for (int k = 0; k < MN; ++k) e = fmax(e, fabs((double)dump[lane * 64 + k] - (double)ref[n * MN + k]));
            if (e < beste) { beste = e; best = n; }
        }
        if (beste < 1e-2) { printf("TMEM lane %3d holds accumulator row n = %2d (max err %.2e)\n", lane, best, beste); ++found; }
    }
    printf("%d of 128 lanes match a reference row (expected 64 for M = 64)\n", found);
    return found == 64 ? 0 : 2;
}
