// Hardware self-test of the tcgen05 building block used by the fused kernels: D[128x64] = A[128x64]·W[64x64]^T
// with the 3xTF32 split, A through TMEM (variant 0/1) or shared memory (variant 2/3), B through the no-swizzle
// K-major shared-memory descriptor.  Exposed through the C ABI so tests/test_gpu_parity.py can pin the
// descriptor encodings on real hardware.
#include <cuda.h>

#include "common.cuh"
#include <cuda_fp16.h>

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

// 16-bit split self-test: A = A1(bf16) + A2(fp16), W = W1(bf16) + W2(fp16); D = A2W2 + A2W1 + A1W2 + A1W1 in
// kind::f16 (variant 8: mixed formats as described; variant 9: everything fp16 with 3 products).
__global__ void __launch_bounds__(128, 1) umma_selftest16_kernel(const float* __restrict__ A,
                                                                 const float* __restrict__ W,
                                                                 float* __restrict__ D, int variant) {
    using namespace umma;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint16_t* B1 = reinterpret_cast<uint16_t*>(smem_raw);        // 8 KB: 64 x 64 x 2 B
    uint16_t* B2 = B1 + 4096;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool all_f16 = (variant == 9);
    // B element (n,k): (k/8)*1024 B + (n/8)*128 B + (n%8)*16 B + (k%8)*2 B
    for (int i = tid; i < 64 * 64; i += 128) {
        const int n = i >> 6, k = i & 63;
        const float w = W[i];
        const uint32_t o = (k >> 3) * 512 + (n >> 3) * 64 + (n & 7) * 8 + (k & 7);
        if (all_f16) {
            const __half h1 = __float2half_rn(w);
            const __half h2 = __float2half_rn(w - __half2float(h1));
            B1[o] = __half_as_ushort(h1);
            B2[o] = __half_as_ushort(h2);
        } else {
            const uint32_t hb = __float_as_uint(w) & 0xFFFF0000u;
            B1[o] = (uint16_t)(hb >> 16);
            B2[o] = __half_as_ushort(__float2half_rn(w - __uint_as_float(hb)));
        }
    }
    if (tid == 0) {
        mbar_init(&mbar, 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(&tmem_base_s, 128);
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_addr = tbase + ((uint32_t)(32 * warp) << 16);
    const uint32_t colA1 = 0, colA2 = 32, colD = 64;
    const float* arow = A + (size_t)tid * 64;
#pragma unroll
    for (int c = 0; c < 2; ++c) {        // 32 elements -> 16 packed columns per chunk
        uint32_t p1[16], p2[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float x0 = arow[32 * c + 2 * j], x1 = arow[32 * c + 2 * j + 1];
            if (all_f16) {
                const __half a0 = __float2half_rn(x0), a1 = __float2half_rn(x1);
                p1[j] = (uint32_t)__half_as_ushort(a0) | ((uint32_t)__half_as_ushort(a1) << 16);
                p2[j] = pack_f16(x0 - __half2float(a0), x1 - __half2float(a1));
            } else {
                uint32_t h0, h1;
                float r0, r1;
                split_bf16_f16(x0, h0, r0);
                split_bf16_f16(x1, h1, r1);
                p1[j] = pack_bf16_trunc(h0, h1);
                p2[j] = pack_f16(r0, r1);
            }
        }
        tmem_st16(lane_addr + colA1 + 16 * c, p1);
        tmem_st16(lane_addr + colA2 + 16 * c, p2);
    }
    wait_st();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        fence_after_sync();
        const int f1 = all_f16 ? 0 : 1;   // format of the leading terms: bf16 (1) or f16 (0)
        const uint64_t d1 = make_b_desc(smem_u32(B1), 1024, 128), d2 = make_b_desc(smem_u32(B2), 1024, 128);
        uint32_t acc = 0;
        auto gemm = [&](uint32_t acol, uint64_t bd, int af, int bf) {
            const uint32_t idesc = make_idesc_f16(128, 64, af, bf);
            for (int ks = 0; ks < 4; ++ks) {      // K = 16 per MMA: 8 packed A columns, 2 K-chunks of B
                mma_f16_ts(tbase + colD, tbase + acol + 8 * ks, bd + (uint64_t)ks * ((2 * 1024) >> 4), idesc, acc);
                acc = 1;
            }
        };
        if (!all_f16) gemm(colA2, d2, 0, 0);      // A2·W2 (f16 x f16)
        gemm(colA2, d1, 0, f1);                   // A2·W1
        gemm(colA1, d2, f1, 0);                   // A1·W2
        gemm(colA1, d1, f1, f1);                  // A1·W1
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
    if (warp == 0) tmem_dealloc(tbase, 128);
}

}  // namespace degnn

extern "C" int distegnn_selftest_umma(const float* A, const float* W,
                                                                              float* D, int variant,
                                                                              void* stream) {
    using namespace degnn;
    DEGNN_CHECK_ARG(A && W && D, "null pointer");
    if (variant >= 8) {
        umma_selftest16_kernel<<<1, 128, 2 * 8192, (cudaStream_t)stream>>>(A, W, D, variant);
        DEGNN_CHECK_LAUNCH();
        return DISTEGNN_OK;
    }
    const int smem = 2 * 16384 + 2 * 32768;
    cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, W, D, variant);
    DEGNN_CHECK_LAUNCH();
    return DISTEGNN_OK;
}

// ---- TMA gather4 building-block self-test ---------------------------------------------------------------------------
// out[g][4][box_floats] = the 4 rows idx[4g..4g+3] of src [n_rows][64] fetched by ONE cp.async.bulk.tensor ...tile::gather4
// each, through a tensor map whose box is WIDER than the row (box_floats = 72 > 64): the out-of-bounds tail must come back
// as zeros and the rows must land at a pitch of box_floats — the layout the edge kernel's staging buffer relies on.
namespace degnn {
__global__ void __launch_bounds__(128) gather4_selftest_kernel(const __grid_constant__ CUtensorMap tm, const int32_t* idx,
                                                               int n_groups, int box_floats, float* out) {
    using namespace umma;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    float* buf = reinterpret_cast<float*>(smem_raw);
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    for (int i = tid; i < n_groups * 4 * box_floats; i += blockDim.x) buf[i] = -7.0f;      // sentinel
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid < 32) {
        if (tid == 0) mbar_expect_tx(&bar, (uint32_t)(n_groups * 4 * box_floats * 4));
        __syncwarp();
        if (elect_one()) {
            for (int g = 0; g < n_groups; ++g) {
                const int4 r = *reinterpret_cast<const int4*>(idx + 4 * g);
                tma_gather4(buf + g * 4 * box_floats, &tm, 0, r.x, r.y, r.z, r.w, &bar);
            }
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    __syncthreads();
    for (int i = tid; i < n_groups * 4 * box_floats; i += blockDim.x) out[i] = buf[i];
}
}  // namespace degnn

extern "C" int distegnn_selftest_gather4(const float* src, int64_t n_rows, const int32_t* idx, int n_groups,
                                         int box_floats, int box_rows, float* out, void* stream) {
    using namespace degnn;
    DEGNN_CHECK_ARG(src && idx && out && n_groups >= 1 && n_groups <= 32, "bad argument");
    DEGNN_CHECK_ARG(box_floats >= 64 && box_floats <= 256 && box_floats % 4 == 0, "bad box width");
    CUtensorMap tm;
    if (int rc = make_rows_tmap(&tm, src, n_rows, 64, box_floats, box_rows)) return rc;
    const int smem = n_groups * 4 * box_floats * 4;
    ensure_dynamic_smem((const void*)gather4_selftest_kernel, smem);
    gather4_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(tm, idx, n_groups, box_floats, out);
    DEGNN_CHECK_LAUNCH();
    return DISTEGNN_OK;
}
