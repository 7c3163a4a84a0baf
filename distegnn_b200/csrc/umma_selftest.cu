// Hardware self-test of the tcgen05 building block used by the fused kernels: D[128x64] = A[128x64]·W[64x64]^T
// with the 3xTF32 split, A through TMEM (variant 0/1) or shared memory (variant 2/3), B through the no-swizzle
// K-major shared-memory descriptor.  Exposed through the C ABI so tests/test_gpu_parity.py can pin the
// descriptor encodings on real hardware.
#include "common.cuh"
#include "umma.cuh"

namespace degnn {

// variant bit0: swap LBO/SBO roles in the descriptors (diagnostic), bit1: A from shared memory,
// bit2: single-pass TF32 (no split) to measure what plain TF32 would cost in accuracy
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const float* __restrict__ A,
                                                               const float* __restrict__ W,
                                                               float* __restrict__ D, int variant) {
    using namespace umma;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    float* Bhi = reinterpret_cast<float*>(smem_raw);              // 16 KB
    float* Blo = Bhi + 4096;                                      // 16 KB
    float* Ahi = Blo + 4096;                                      // 32 KB (SS variants)
    float* Alo = Ahi + 8192;                                      // 32 KB
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool swap = variant & 1, a_smem = variant & 2, single = variant & 4;

    // stage B (hi/lo) in the canonical layout
    for (int i = tid; i < 64 * 64; i += 128) {
        int n = i >> 6, k = i & 63;
        uint32_t hi, lo;
        split_tf32(W[i], hi, lo);
        if (single) { hi = __float_as_uint(W[i]); lo = 0; }
        Bhi[b_elem_offset(n, k)] = __uint_as_float(hi);
        Blo[b_elem_offset(n, k)] = __uint_as_float(lo);
    }
    // A for the SS variants: same core-matrix layout with 16 row groups: (k/4)*2048 + (m/8)*128 + (m%8)*16 + (k%4)*4
    if (a_smem) {
        for (int i = tid; i < 128 * 64; i += 128) {
            int m = i >> 6, k = i & 63;
            uint32_t hi, lo;
            split_tf32(A[i], hi, lo);
            if (single) { hi = __float_as_uint(A[i]); lo = 0; }
            uint32_t off = (k >> 2) * 512 + (m >> 3) * 32 + (m & 7) * 4 + (k & 3);
            Ahi[off] = __uint_as_float(hi);
            Alo[off] = __uint_as_float(lo);
        }
    }
    if (tid == 0) {
        mbar_init(&mbar, 1);
        fence_mbar_init();
    }
    __syncwarp();   // .sync.aligned below needs the whole warp converged
    if (warp == 0) tmem_alloc(&tmem_base_s, 256);
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_addr = tbase + ((uint32_t)(32 * warp) << 16);
    const uint32_t colAhi = 0, colAlo = 64, colD = 128;

    if (!a_smem) {
        // thread = row: write hi/lo of its 64 values into TMEM
        const float* arow = A + (size_t)tid * 64;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float v = arow[16 * c + j];
                split_tf32(v, hi[j], lo[j]);
                if (single) { hi[j] = __float_as_uint(v); lo[j] = 0; }
            }
            tmem_st16(lane_addr + colAhi + 16 * c, hi);
            tmem_st16(lane_addr + colAlo + 16 * c, lo);
        }
        wait_st();
    }
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        fence_after_sync();
        const uint32_t idesc = make_idesc_tf32(128, 64);
        const uint32_t lboB = swap ? B_SBO : B_LBO, sboB = swap ? B_LBO : B_SBO;
        const uint32_t lboA = swap ? 128u : 2048u, sboA = swap ? 2048u : 128u;
        uint32_t acc = 0;
        // order: small terms first
        for (int pass = 0; pass < (single ? 1 : 3); ++pass) {
            const bool a_lo = (!single && pass == 0), b_lo = (!single && pass == 1);
            const float* Bs = b_lo ? Blo : Bhi;
            const float* As = a_lo ? Alo : Ahi;
            for (int ks = 0; ks < 8; ++ks) {
                uint64_t bd = make_b_desc(smem_u32(Bs) + ks * 2 * B_LBO, lboB, sboB);
                if (a_smem) {
                    uint64_t ad = make_b_desc(smem_u32(As) + ks * 2 * 2048, lboA, sboA);
                    mma_tf32_ss(tbase + colD, ad, bd, idesc, acc);
                } else {
                    mma_tf32_ts(tbase + colD, tbase + (a_lo ? colAlo : colAhi) + 8 * ks, bd, idesc, acc);
                }
                acc = 1;
            }
        }
        mma_commit(&mbar);
    }
    __syncwarp();
    mbar_wait(&mbar, 0);
    __syncwarp();
    fence_after_sync();
    float* drow = D + (size_t)tid * 64;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t r[16];
        tmem_ld16(lane_addr + colD + 16 * c, r);
        wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) drow[16 * c + j] = __uint_as_float(r[j]);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 256);
}

}  // namespace degnn

extern "C" int distegnn_selftest_umma(const float* A, const float* W,
                                                                              float* D, int variant,
                                                                              void* stream) {
    using namespace degnn;
    DEGNN_CHECK_ARG(A && W && D, "null pointer");
    const int smem = 2 * 16384 + 2 * 32768;
    cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, W, D, variant);
    DEGNN_CHECK_LAUNCH();
    return DISTEGNN_OK;
}
