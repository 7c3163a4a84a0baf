// tcgen05 / TMEM / mbarrier / bulk-copy PTX wrappers for sm_100a (hand-written; encodings follow the PTX
// ISA "tcgen05" chapter: shared-memory matrix descriptor, instruction descriptor kind::tf32).
//
// Everything here serves one pattern: a 128-row tile whose rows are owned by 128 threads (thread r of a
// warpgroup <-> TMEM lane r), the A operand written to TMEM by those threads (tcgen05.st 32x32b), the
// 64x64 weight matrix B resident in shared memory in the canonical no-swizzle K-major layout, and the
// fp32 accumulator read back row-per-thread (tcgen05.ld 32x32b).  fp32 accuracy comes from the 3xTF32
// split  A·W ≈ A_lo·W_hi + A_hi·W_lo + A_hi·W_hi  (hi = top 19 bits, lo = exact remainder).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace degnn {
namespace umma {

// ---- shared-memory layout of a 64(N) x 64(K) fp32 B operand, K-major, no swizzle --------------------
// Core matrix = 8 rows x 16 bytes (4 tf32) stored contiguously (128 B).  Address of element (n,k):
//   (k/4)*LBO + (n/8)*SBO + (n%8)*16 + (k%4)*4,   LBO = 1024 B (next K chunk), SBO = 128 B (next 8 rows)
// One tcgen05.mma kind::tf32 consumes K=8 = two K chunks: descriptor start = base + kstep*2*LBO.
constexpr uint32_t B_LBO = 1024, B_SBO = 128, B_BYTES = 16 * 1024;

__device__ __forceinline__ uint32_t b_elem_offset(int n, int k) {   // in floats
    return (uint32_t)((k >> 2) * (B_LBO / 4) + (n >> 3) * (B_SBO / 4) + (n & 7) * 4 + (k & 3));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// 64-bit shared-memory matrix descriptor (PTX ISA, tcgen05 "matrix-descriptor"):
//   [0,14) start address >> 4   [16,30) leading-dim byte offset >> 4   [32,46) stride-dim byte offset >> 4
//   [46,48) = 0b01 (sm_100 descriptor version)   [49,52) base offset = 0   [61,64) swizzle mode = 0 (none)
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

// 32-bit instruction descriptor for kind::tf32, fp32 accumulate, A and B K-major, dense:
//   [4,6) D format = 1 (f32)  [7,10) A format = 2 (tf32)  [10,13) B format = 2 (tf32)
//   [15] A major = 0 (K)  [16] B major = 0 (K)  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::f16 variant: formats 0 = f16, 1 = bf16 (A at [7,10), B at [10,13)), fp32 accumulate, K = 16 per MMA
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_fmt, int b_fmt) {
    return (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 2-term 16-bit split of an fp32 value: x ≈ bf16(x) [truncated: top 16 bits] + fp16(x − bf16(x))
// (8 + 11 significant bits, fp32 exponent range for the leading term)
__device__ __forceinline__ void split_bf16_f16(float x, uint32_t& hi_bf16_bits, float& rem) {
    hi_bf16_bits = __float_as_uint(x) & 0xFFFF0000u;
    rem = x - __uint_as_float(hi_bf16_bits);
}
__device__ __forceinline__ uint32_t pack_bf16_trunc(uint32_t even_bits, uint32_t odd_bits) {
    return (even_bits >> 16) | (odd_bits & 0xFFFF0000u);      // element k even in the low half
}
__device__ __forceinline__ uint32_t pack_f16(float even, float odd) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(odd), "f"(even));   // first source -> upper half
    return r;
}

// ---- TMEM allocation (one warp, whole warp executes) -----------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- fences / waits ----------------------------------------------------------------------------------
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// make generic-proxy smem writes (st.shared) visible to the async proxy (tensor core / bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM <-> registers, 32 lanes x 32 bit, 16 consecutive columns per call --------------------------
// Warp w of a warpgroup may only touch lanes 32*(w%4) .. +31; taddr = (lane_base << 16) | column.
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
// tcgen05.wait::ld that also "produces" the 8 registers of an earlier tmem_ld8: consumers of r cannot be scheduled above
// the wait, which makes it safe to issue the NEXT chunk's load before computing on this one (software pipelining of the
// TMEM reads: 64 B/clk/SM and ~ a dozen cycles of latency per load).
__device__ __forceinline__ void wait_ld8(uint32_t (&r)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
                 :
                 : "memory");
}
__device__ __forceinline__ void wait_ld16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// ---- MMA: D[tmem] (+)= A[tmem] · B[smem]^T, M=128, N=64, K=8 (tf32), issued by ONE thread -------------
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same with A from shared memory (descriptor) — used by the self-test to cross-check both operand paths
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void mma_commit(uint64_t* mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar))
                 : "memory");
}

// ---- mbarrier ----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// UMMA_WAIT_HINT_NS > 0: pass a suspend-time hint to try_wait — ptxas then emits TRYWAIT / NANOSLEEP.SYNCS / PHASECHK, the
// warp sleeps until the phase completes (or the hint expires) instead of re-polling and stealing issue slots from the
// warps that have work (polls were ~5 % of the edge kernel's instructions).
#ifndef UMMA_WAIT_HINT_NS
#define UMMA_WAIT_HINT_NS 0
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#if UMMA_WAIT_HINT_NS > 0
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"((uint32_t)UMMA_WAIT_HINT_NS)
        : "memory");
#else
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
#endif
}

// ---- bulk async copy global -> shared (TMA engine, no tensor map), completes on an mbarrier ----------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// One lane of the (converged) warp, chosen by the hardware.  Code under `if (elect_one())` is known to ptxas to run in a
// single lane, so per-thread values feeding UBLKCP / UTCHMMA operands move to uniform registers with plain R2URs — no
// ELECT / R2UR / PLOP3 / BRA.ANY loop over the active lanes (9 instructions per bulk copy otherwise).
__device__ __forceinline__ bool elect_one() {
    uint32_t e;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(e));
    return e != 0;
}

// ---- TMA row gather: FOUR rows of a 2-D tensor (tensor map) per instruction into 4 consecutive box rows -------------
__device__ __forceinline__ void tma_gather4(void* smem_dst, const void* tmap, int c0, int r0, int r1, int r2, int r3,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, "
        "%5, %6}], [%7];" ::"r"(smem_u32(smem_dst)),
        "l"(tmap), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
        : "memory");
}

// ---- bulk async copy shared -> global (TMA engine), tracked by the issuing thread's bulk async-group --------------
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- per-thread async copies global -> shared (LDGSTS: per-lane addresses, no registers, generic proxy) ------------
// Completion is per issuing thread (commit_group / wait_group); a thread that only reads what it copied itself needs
// no barrier.  16-byte form bypasses L1 (.cg); the 4/8-byte forms allocate in L1 (.ca is the only variant).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// two adjacent fp32 added to global memory with one reduction (sm_90+), 8-byte aligned
__device__ __forceinline__ void red_add_v2(float* gdst, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(gdst), "f"(a), "f"(b) : "memory");
}

// named barrier among `nthreads` threads (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// 3xTF32 split: hi keeps sign, exponent and the top 10 mantissa bits; lo = x - hi is exact in fp32
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

}  // namespace umma
}  // namespace degnn
