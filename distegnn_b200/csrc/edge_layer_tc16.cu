// Real<->real edge stage on the 5th-gen tensor cores — production kernel behind distegnn_edge_layer_fwd.
// Replaces reference models/FastEGNN.py:237-246 (coord2radial), 144-150 (edge_model), 169-177 (edge part of
// coord_model_vel), 206 (edge part of node_model) and the scatter_add_ of :322-337 (twins models/basic.py:22-66).
//
// Numerics: the two 64x64 layers of every edge run as tile GEMMs D[128x64] = A[128x64]·Wᵀ in kind::f16 with a
// 2-term fp16 split of BOTH operands:  x = hi + lo, hi = fp16(x), lo = fp16(x − hi)  (22 significant bits),
// D = lo·Whi + hi·Wlo + hi·Whi accumulated in fp32 — measured error of the building block 8.5e-7 (torch fp32
// 1.6e-6, 3xTF32 2.4e-6, plain TF32 3.5e-3; scripts/umma_selftest.py).  fp16 range is handled per row: a row
// whose largest activation exceeds 3e4 is re-encoded with a power-of-two scale s and its accumulator is
// multiplied by 1/s in the epilogue (exact), so the result is range-safe like fp32.
//
// One CTA per SM, 512 threads = 4 independent tile groups of 4 warps; thread r of a group owns edge r of the
// group's current 128-edge tile end to end (TMEM lane r).  TMEM per group: A_hi 32 + A_lo 32 + D 64 columns.
//   stage 0  TMA bulk copies (cp.async.bulk, one 256-B row per edge) stage the neighbour rows Q[col] of the
//            next tile in shared memory; completion on an mbarrier;
//   stage 1  a1 = SiLU(P[row] + Q[col] + w_r·r + W_e·a) -> fp16 hi/lo -> tcgen05.st;
//   MMA 1    12 x tcgen05.mma (M128 N64 K16) -> D;      stage 2  m = SiLU(D + b2): row to shared (segment sum)
//            and hi/lo to TMEM;   MMA 2 (φ head) overlapped with the segment sum of m over destination rows;
//   stage 3  φ = w3·SiLU(D + bc); Δx·φ reduced over runs of equal row by warp shuffles; RED.ADD.
// With four tiles in flight per SM one group's MMA / barrier / memory waits are covered by the others.
#include <cuda_fp16.h>

#include "common.cuh"
#include "umma.cuh"

namespace degnn {

struct EdgeT16Args {
    int64_t N, E;
    int A;
    unsigned flags;
    const int32_t* row;
    const int32_t* col;
    const float* ea;
    const float* x4;
    const float* P;
    const float* Q;
    const float* w1r;
    const float* w1e;
    const float* w2;   // k-major [k][n]
    const float* b2;
    const float* wc;   // k-major [k][n]
    const float* bc;
    const float* w3;
    float* agg_m;
    float* agg_x;
};

constexpr int T16_THREADS = 512, T16_GROUPS = 4, T16_GROUP = 128;
constexpr int T16_QROW = 68;
constexpr int T16_QBUF = TILE_M * T16_QROW;
constexpr int T16_W_HALFS = 64 * 64;                      // fp16 elements per weight matrix (8 KB)
constexpr int T16_SMEM_BYTES = 4 * T16_W_HALFS * 2        // W2 hi/lo, Wc hi/lo
                               + T16_GROUPS * T16_QBUF * 4
                               + (4 * H + DISTEGNN_MAX_EDGE_ATTR * H) * 4   // b2, bc, w3, w1r, w1e
                               + T16_GROUPS * TILE_M * 4  // srow
                               + T16_GROUPS * 4 * 4       // run-start bit masks (one word per warp)
                               + 128;                     // mbarriers + tmem base
constexpr uint32_t T16_LBO = 1024, T16_SBO = 128;         // fp16 K-major no-swizzle: 8 rows x 8 halfs per core matrix
constexpr float T16_RANGE = 3.0e4f;
#ifndef T16_CHUNK_UNROLL
#define T16_CHUNK_UNROLL 2     // chunks (of 16 columns) unrolled per stage loop: trades code size (I-cache) for ILP
#endif
constexpr int kChunkUnroll = T16_CHUNK_UNROLL;

// weight W[n][k] = wt_kmajor[k*64+n] -> fp16 hi/lo at (k/8)*512 + (n/8)*64 + (n%8)*8 + k%8 (in halfs)
__device__ __forceinline__ void stage_weight_f16(__half* hi, __half* lo, const float* __restrict__ wt_kmajor, int tid,
                                                 int nthreads) {
    for (int i = tid; i < H * H; i += nthreads) {
        const int k = i >> 6, n = i & 63;
        const float w = __ldg(wt_kmajor + i);
        const __half h = __float2half_rn(w);
        const uint32_t o = (uint32_t)(k >> 3) * 512u + (uint32_t)(n >> 3) * 64u + (uint32_t)(n & 7) * 8u + (k & 7);
        hi[o] = h;
        lo[o] = __float2half_rn(w - __half2float(h));
    }
}

// 12 MMAs: D = Alo·Bhiᵀ + Ahi·Bloᵀ + Ahi·Bhiᵀ (K = 64 = 4 steps of 16), then commit to `bar`
__device__ __forceinline__ void issue_gemm_f16x3(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                                 uint32_t idesc, uint64_t* bar) {
    constexpr uint64_t KSTEP = (2 * T16_LBO) >> 4;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) umma::mma_f16_ts(d, a_lo + 8 * ks, b_hi + ks * KSTEP, idesc, ks > 0);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) umma::mma_f16_ts(d, a_hi + 8 * ks, b_lo + ks * KSTEP, idesc, 1);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) umma::mma_f16_ts(d, a_hi + 8 * ks, b_hi + ks * KSTEP, idesc, 1);
    umma::mma_commit(bar);
}

// 16 fp32 values (·s) -> 8 packed hi words + 8 packed lo words; `mx` tracks the running max of the hi halves
template <bool SCALED>
__device__ __forceinline__ void split16(const float (&v)[16], float s, uint32_t (&hi)[8], uint32_t (&lo)[8], __half2& mx) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float x0 = SCALED ? v[2 * j] * s : v[2 * j], x1 = SCALED ? v[2 * j + 1] * s : v[2 * j + 1];
        const __half2 h = __floats2half2_rn(x0, x1);            // x0 -> low half (even k)
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
        mx = __hmax2(mx, h);
        hi[j] = *reinterpret_cast<const uint32_t*>(&h);
        lo[j] = *reinterpret_cast<const uint32_t*>(&l);
    }
}
__device__ __forceinline__ bool row_overflow(__half2 mx) {
    return fmaxf(__low2float(mx), __high2float(mx)) > T16_RANGE;
}
// power-of-two scale that brings `rowmax` below 2^15, and its inverse
__device__ __forceinline__ void range_scale(float rowmax, float& s, float& inv_s) {
    const uint32_t eb = (__float_as_uint(rowmax) >> 23) & 0xffu;       // biased exponent
    const uint32_t sb = eb > 141u ? 268u - eb : 127u;                   // rowmax >= 2^15 -> s = 2^(14-e)
    s = __uint_as_float((sb < 1u ? 1u : sb) << 23);
    inv_s = 1.0f / s;
}

template <int AT>
__global__ void __launch_bounds__(T16_THREADS, 1) edge_layer_t16_kernel(const EdgeT16Args a) {
    using namespace umma;
    constexpr int AMAX = AT >= 0 ? (AT > 0 ? AT : 1) : DISTEGNN_MAX_EDGE_ATTR;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __half* W2hi = reinterpret_cast<__half*>(smem_raw);
    __half* W2lo = W2hi + T16_W_HALFS;
    __half* Wchi = W2lo + T16_W_HALFS;
    __half* Wclo = Wchi + T16_W_HALFS;
    float* qbufs = reinterpret_cast<float*>(Wclo + T16_W_HALFS);     // [4][QBUF]
    float* b2s = qbufs + T16_GROUPS * T16_QBUF;
    float* bcs = b2s + H;
    float* w3s = bcs + H;
    float* w1rs = w3s + H;
    float* w1es = w1rs + H;
    int* srow_all = reinterpret_cast<int*>(w1es + DISTEGNN_MAX_EDGE_ATTR * H);   // [4][128]
    uint32_t* rmask_all = reinterpret_cast<uint32_t*>(srow_all + T16_GROUPS * TILE_M);   // [4][4]
    uint64_t* bars = reinterpret_cast<uint64_t*>(rmask_all + T16_GROUPS * 4);      // [4][2]
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 2 * T16_GROUPS);

    const int tid = threadIdx.x;
    const int grp = tid >> 7;              // tile group 0..3
    const int t = tid & 127;               // edge (row) of the tile owned by this thread
    const int lane = tid & 31;
    const int wq = (tid >> 5) & 3;         // warp inside the group == TMEM lane quarter
    const int A = AT >= 0 ? AT : a.A;
    const bool normalize = a.flags & DISTEGNN_FLAG_NORMALIZE;
    const bool need_m = !(a.flags & DISTEGNN_FLAG_LAST);

    // ---- one-time setup -------------------------------------------------------------------------
    stage_weight_f16(W2hi, W2lo, a.w2, tid, T16_THREADS);
    stage_weight_f16(Wchi, Wclo, a.wc, tid, T16_THREADS);
    if (tid < H) {
        b2s[tid] = a.b2[tid];
        bcs[tid] = a.bc[tid];
        w3s[tid] = a.w3[tid];
        w1rs[tid] = a.w1r[tid];
    }
    for (int i = tid; i < A * H; i += T16_THREADS) w1es[i] = a.w1e[i];
    if (tid == 0) {
        for (int i = 0; i < 2 * T16_GROUPS; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    __syncwarp();
    if ((tid >> 5) == 0) tmem_alloc(tmem_base_s, 512);
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();

    const uint32_t tbase = *tmem_base_s;
    const uint32_t col0 = tbase + (uint32_t)grp * 128u;
    const uint32_t tA_hi = col0, tA_lo = col0 + 32, tD = col0 + 64;
    const uint32_t lane_off = ((uint32_t)(32 * wq)) << 16;
    const uint32_t idesc = make_idesc_f16(128, 64, 0, 0);
    const uint64_t dW2hi = make_b_desc(smem_u32(W2hi), T16_LBO, T16_SBO), dW2lo = make_b_desc(smem_u32(W2lo), T16_LBO, T16_SBO);
    const uint64_t dWchi = make_b_desc(smem_u32(Wchi), T16_LBO, T16_SBO), dWclo = make_b_desc(smem_u32(Wclo), T16_LBO, T16_SBO);
    float* qb = qbufs + grp * T16_QBUF;
    float* myq = qb + t * T16_QROW;
    int* srow = srow_all + grp * TILE_M;
    uint32_t* rmask = rmask_all + grp * 4;
    uint64_t* qbar = bars + grp * 2;
    uint64_t* mbar = bars + grp * 2 + 1;
    const uint32_t bar_id = 1 + grp;

    const int64_t num_tiles = (a.E + TILE_M - 1) / TILE_M;
    const int64_t stride = (int64_t)gridDim.x * T16_GROUPS;
    int64_t tile = (int64_t)blockIdx.x * T16_GROUPS + grp;

    int row_c = -1, rr_c = 0;
    float dx = 0.f, dy = 0.f, dz = 0.f, radial = 0.f;
    float ea_c[AMAX];
    auto load_edge = [&](int64_t tl, int& r, int& rr, int& c, float (&ea)[AMAX]) {
        const int64_t e = tl * TILE_M + t;
        r = -1;
        rr = 0;
        c = 0;
#pragma unroll
        for (int k = 0; k < AMAX; ++k) ea[k] = 0.f;
        if (tl < num_tiles && e < a.E) {
            r = __ldg(a.row + e);
            c = __ldg(a.col + e);
            rr = r;
#pragma unroll
            for (int k = 0; k < AMAX; ++k)
                if (k < A) ea[k] = __ldg(a.ea + e * A + k);
        }
    };
    auto prefetch_q = [&](int64_t tl, int r, int c) {
        if (tl < num_tiles) {
            if (t == 0) {
                const int64_t nvalid = min((int64_t)TILE_M, a.E - tl * TILE_M);
                mbar_expect_tx(qbar, (uint32_t)nvalid * (H * 4));
            }
            if (r >= 0) bulk_g2s(myq, a.Q + (size_t)c * H, H * 4, qbar);
        }
    };

    if (tile < num_tiles) {
        int col_c;
        load_edge(tile, row_c, rr_c, col_c, ea_c);
        prefetch_q(tile, row_c, col_c);
        const float4 xi = ldg4(a.x4 + (size_t)rr_c * 4), xj = ldg4(a.x4 + (size_t)col_c * 4);
        dx = xi.x - xj.x; dy = xi.y - xj.y; dz = xi.z - xj.z;
        radial = dx * dx + dy * dy + dz * dz;
        if (normalize) {
            const float inv = 1.0f / (sqrtf(radial) + 1e-8f);
            dx *= inv; dy *= inv; dz *= inv;
        }
    }

    for (int it = 0; tile < num_tiles; ++it, tile += stride) {
        // ---- stage 1: a1 = SiLU(P_i + Q_j + w_r·r + W_e·a) -> fp16 hi/lo -> TMEM ---------------------------
        mbar_wait(qbar, (uint32_t)(it & 1));
        __syncwarp();
        const float* prow = a.P + (size_t)rr_c * H;
        auto pre_chunk = [&](int c, float (&v)[16]) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const int cc = 16 * c + 4 * j4;
                float4 pre = fma4(radial, *reinterpret_cast<const float4*>(w1rs + cc),
                                  add4(ldg4(prow + cc), *reinterpret_cast<const float4*>(myq + cc)));
#pragma unroll
                for (int k = 0; k < AMAX; ++k)
                    if (k < A) pre = fma4(ea_c[k], *reinterpret_cast<const float4*>(w1es + k * H + cc), pre);
                pre = silu4(pre);
                v[4 * j4 + 0] = pre.x; v[4 * j4 + 1] = pre.y; v[4 * j4 + 2] = pre.z; v[4 * j4 + 3] = pre.w;
            }
        };
        float inv_s1 = 1.0f;
        {
            __half2 mx = __floats2half2_rn(0.f, 0.f);
#pragma unroll kChunkUnroll
            for (int c = 0; c < 4; ++c) {
                float v[16];
                uint32_t hi[8], lo[8];
                pre_chunk(c, v);
                split16<false>(v, 1.0f, hi, lo, mx);
                tmem_st8(lane_off + tA_hi + 8 * c, hi);
                tmem_st8(lane_off + tA_lo + 8 * c, lo);
            }
            if (__any_sync(FULL, row_overflow(mx))) {      // cold: some row of this warp leaves the fp16 range
                float fm = 0.f, sc;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    float v[16];
                    pre_chunk(c, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) fm = fmaxf(fm, v[j]);
                }
                range_scale(fm, sc, inv_s1);
                wait_st();
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    float v[16];
                    uint32_t hi[8], lo[8];
                    pre_chunk(c, v);
                    split16<true>(v, sc, hi, lo, mx);
                    tmem_st8(lane_off + tA_hi + 8 * c, hi);
                    tmem_st8(lane_off + tA_lo + 8 * c, lo);
                }
            }
        }
        wait_st();
        srow[t] = row_c;
        fence_before_sync();
        named_bar(bar_id, T16_GROUP);      // A complete; D of the previous tile fully read by the whole group

        // ---- MMA 1; meanwhile read the next tile's edge list and coordinates -------------------------------
        if (t == 0) {
            fence_after_sync();
            issue_gemm_f16x3(tD, tA_hi, tA_lo, dW2hi, dW2lo, idesc, mbar);
        }
        __syncwarp();
        {   // bit i of rmask[w] = edge 32w+i starts a new run of equal destination rows (read after barrier 2)
            const int prev = t > 0 ? srow[t - 1] : -2;
            const uint32_t starts = __ballot_sync(FULL, prev != row_c);
            if (lane == 0) rmask[wq] = starts;
        }
        int row_n, rr_n, col_n;
        float ea_n[AMAX];
        load_edge(tile + stride, row_n, rr_n, col_n, ea_n);
        prefetch_l1(a.P + (size_t)rr_n * H);
        prefetch_l1(a.P + (size_t)rr_n * H + 32);
        const float4 xi_n = ldg4(a.x4 + (size_t)rr_n * 4), xj_n = ldg4(a.x4 + (size_t)col_n * 4);

        mbar_wait(mbar, 0);
        __syncwarp();
        fence_after_sync();

        // ---- stage 2: m = SiLU(D/s + b2) -> row to shared (segment sum) and fp16 hi/lo -> TMEM ---------------
        auto m_chunk = [&](int c, float (&v)[16], bool store) {
            uint32_t d[16];
            tmem_ld16(lane_off + tD + 16 * c, d);
            wait_ld();
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const int cc = 16 * c + 4 * j4;
                const float4 bb = *reinterpret_cast<const float4*>(b2s + cc);
                float4 m;
                m.x = silu(fmaf(__uint_as_float(d[4 * j4 + 0]), inv_s1, bb.x));
                m.y = silu(fmaf(__uint_as_float(d[4 * j4 + 1]), inv_s1, bb.y));
                m.z = silu(fmaf(__uint_as_float(d[4 * j4 + 2]), inv_s1, bb.z));
                m.w = silu(fmaf(__uint_as_float(d[4 * j4 + 3]), inv_s1, bb.w));
                if (store) *reinterpret_cast<float4*>(myq + cc) = m;
                v[4 * j4 + 0] = m.x; v[4 * j4 + 1] = m.y; v[4 * j4 + 2] = m.z; v[4 * j4 + 3] = m.w;
            }
        };
        float inv_s2 = 1.0f;
        {
            __half2 mx = __floats2half2_rn(0.f, 0.f);
#pragma unroll kChunkUnroll
            for (int c = 0; c < 4; ++c) {
                float v[16];
                uint32_t hi[8], lo[8];
                m_chunk(c, v, need_m);
                split16<false>(v, 1.0f, hi, lo, mx);
                tmem_st8(lane_off + tA_hi + 8 * c, hi);
                tmem_st8(lane_off + tA_lo + 8 * c, lo);
            }
            if (__any_sync(FULL, row_overflow(mx))) {      // cold
                float fm = 0.f, sc;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    float v[16];
                    m_chunk(c, v, false);
#pragma unroll
                    for (int j = 0; j < 16; ++j) fm = fmaxf(fm, v[j]);
                }
                range_scale(fm, sc, inv_s2);
                wait_st();
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    float v[16];
                    uint32_t hi[8], lo[8];
                    m_chunk(c, v, false);
                    split16<true>(v, sc, hi, lo, mx);
                    tmem_st8(lane_off + tA_hi + 8 * c, hi);
                    tmem_st8(lane_off + tA_lo + 8 * c, lo);
                }
            }
        }
        wait_st();
        fence_before_sync();
        named_bar(bar_id, T16_GROUP);      // m tile visible in shared, A complete, D fully read

        // ---- MMA 2 (φ head) overlapped with the segment sum of m --------------------------------------------
        if (t == 0) {
            fence_after_sync();
            issue_gemm_f16x3(tD, tA_hi, tA_lo, dWchi, dWclo, idesc, mbar);
        }
        __syncwarp();
#ifdef DEGNN_DIAG_NO_SEGSUM
        if (false) {
#else
        if (need_m) {
#endif
            // thread (column c, half of the tile): one RED per (run of equal destination row, column)
            const int c = t & 63, hh = t >> 6, eb = hh * 64;
            const float* colp = qb + eb * T16_QROW + c;
            uint64_t M = ((uint64_t)rmask[2 * hh + 1] << 32) | rmask[2 * hh] | 1ull;   // run starts in this half
            if (__popcll(M) <= 24) {
                // few runs (the usual radius-graph regime): walk run by run, 2 instructions per edge
                while (M) {
                    const int e0 = __ffsll((long long)M) - 1;
                    M &= M - 1;
                    const int e1 = M ? __ffsll((long long)M) - 1 : 64;
                    float s0 = 0.f, s1 = 0.f;
                    int e = e0;
                    for (; e + 1 < e1; e += 2) {
                        s0 += colp[e * T16_QROW];
                        s1 += colp[(e + 1) * T16_QROW];
                    }
                    if (e < e1) s0 += colp[e * T16_QROW];
                    const int r = srow[eb + e0];
                    if (r >= 0) atomicAdd(a.agg_m + (size_t)r * H + c, s0 + s1);
                }
            } else {
                // many short runs (sparse partitions): per-edge walk
                int cur = srow[eb];
                float s0 = 0.f;
#pragma unroll 4
                for (int e = 0; e < 64; ++e) {
                    const int r0 = srow[eb + e];
                    if (r0 != cur) {
                        if (cur >= 0) atomicAdd(a.agg_m + (size_t)cur * H + c, s0);
                        s0 = 0.f;
                        cur = r0;
                    }
                    s0 += colp[e * T16_QROW];
                }
                if (cur >= 0) atomicAdd(a.agg_m + (size_t)cur * H + c, s0);
            }
            fence_proxy_async_smem();          // generic accesses to qb ordered before the TMA refill below
        }
        named_bar(bar_id, T16_GROUP);          // whole group done with the staging buffer
        prefetch_q(tile + stride, row_n, col_n);

        mbar_wait(mbar, 1);
        __syncwarp();
        fence_after_sync();

        // ---- stage 3: φ = w3·SiLU(D/s + bc); Δx·φ summed per destination row --------------------------------
        float ph0 = 0.f, ph1 = 0.f, ph2 = 0.f, ph3 = 0.f;      // four independent FMA chains
#pragma unroll kChunkUnroll
        for (int c = 0; c < 4; ++c) {
            uint32_t d[16];
            tmem_ld16(lane_off + tD + 16 * c, d);
            wait_ld();
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const int cc = 16 * c + 4 * j4;
                const float4 bb = *reinterpret_cast<const float4*>(bcs + cc);
                const float4 ww = *reinterpret_cast<const float4*>(w3s + cc);
                ph0 = fmaf(silu(fmaf(__uint_as_float(d[4 * j4 + 0]), inv_s2, bb.x)), ww.x, ph0);
                ph1 = fmaf(silu(fmaf(__uint_as_float(d[4 * j4 + 1]), inv_s2, bb.y)), ww.y, ph1);
                ph2 = fmaf(silu(fmaf(__uint_as_float(d[4 * j4 + 2]), inv_s2, bb.z)), ww.z, ph2);
                ph3 = fmaf(silu(fmaf(__uint_as_float(d[4 * j4 + 3]), inv_s2, bb.w)), ww.w, ph3);
            }
        }
        const float phi = (ph0 + ph1) + (ph2 + ph3);
        fence_before_sync();                   // D reads ordered before the next tile's MMA 1
        {
            float sx = dx * phi, sy = dy * phi, sz = dz * phi;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int rk = __shfl_up_sync(FULL, row_c, o);
                const float ox = __shfl_up_sync(FULL, sx, o), oy = __shfl_up_sync(FULL, sy, o),
                            oz = __shfl_up_sync(FULL, sz, o);
                if (lane >= o && rk == row_c) { sx += ox; sy += oy; sz += oz; }
            }
            const int rnext = __shfl_down_sync(FULL, row_c, 1);
            if (row_c >= 0 && (lane == 31 || rnext != row_c)) {
                float* dst = a.agg_x + (size_t)row_c * 4;
                atomicAdd(dst + 0, sx);
                atomicAdd(dst + 1, sy);
                atomicAdd(dst + 2, sz);
            }
        }

        // ---- roll the prefetched edge into place --------------------------------------------------------
        row_c = row_n; rr_c = rr_n;
#pragma unroll
        for (int k = 0; k < AMAX; ++k) ea_c[k] = ea_n[k];
        dx = xi_n.x - xj_n.x; dy = xi_n.y - xj_n.y; dz = xi_n.z - xj_n.z;
        radial = dx * dx + dy * dy + dz * dz;
        if (normalize) {
            const float inv = 1.0f / (sqrtf(radial) + 1e-8f);
            dx *= inv; dy *= inv; dz *= inv;
        }
    }

    fence_before_sync();
    __syncthreads();
    if ((tid >> 5) == 0) tmem_dealloc(tbase, 512);
}

}  // namespace degnn

extern "C" int distegnn_edge_layer_fwd(int64_t n_nodes, int64_t n_edges, int A, int C, int Na, unsigned flags,
                                       const int32_t* row, const int32_t* col, const float* edge_attr_sorted,
                                       const float* x4, const float* P, const float* Q,
                                       const float* layer_params, float* agg_m, float* agg_x, void* stream) {
    using namespace degnn;
    if (int rc = check_dims(A, C, Na)) return rc;
    if (n_edges == 0) return DISTEGNN_OK;
    DEGNN_CHECK_ARG(n_nodes > 0 && n_edges > 0, "negative size");
    DEGNN_CHECK_ARG(row && col && x4 && P && Q && layer_params && agg_x, "null pointer");
    DEGNN_CHECK_ARG(A == 0 || edge_attr_sorted, "null edge_attr with edge_attr_nf > 0");
    DEGNN_CHECK_ARG((flags & DISTEGNN_FLAG_LAST) || agg_m, "null agg_m");
    Layout L = make_layout(A, C, Na);
    EdgeT16Args a;
    a.N = n_nodes; a.E = n_edges; a.A = A; a.flags = flags;
    a.row = row; a.col = col; a.ea = edge_attr_sorted; a.x4 = x4; a.P = P; a.Q = Q;
    a.w1r = layer_params + L.off[DISTEGNN_P_E_W1R];
    a.w1e = layer_params + L.off[DISTEGNN_P_E_W1E];
    a.w2 = layer_params + L.off[DISTEGNN_P_E_W2];
    a.b2 = layer_params + L.off[DISTEGNN_P_E_B2];
    a.wc = layer_params + L.off[DISTEGNN_P_E_WC];
    a.bc = layer_params + L.off[DISTEGNN_P_E_BC];
    a.w3 = layer_params + L.off[DISTEGNN_P_E_W3];
    a.agg_m = agg_m; a.agg_x = agg_x;
    const int64_t tiles = (n_edges + TILE_M - 1) / TILE_M;
    int64_t grid = (tiles + T16_GROUPS - 1) / T16_GROUPS;
    if (grid > sm_count()) grid = sm_count();
    auto launch = [&](auto kern) {
        ensure_dynamic_smem((const void*)kern, (int)T16_SMEM_BYTES);
        kern<<<(unsigned)grid, T16_THREADS, T16_SMEM_BYTES, (cudaStream_t)stream>>>(a);
    };
    switch (A) {
        case 0: launch(edge_layer_t16_kernel<0>); break;
        case 1: launch(edge_layer_t16_kernel<1>); break;
        case 2: launch(edge_layer_t16_kernel<2>); break;
        default: launch(edge_layer_t16_kernel<-1>); break;
    }
    DEGNN_CHECK_LAUNCH();
    return DISTEGNN_OK;
}
