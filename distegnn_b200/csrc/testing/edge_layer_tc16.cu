// Real<->real edge stage on the 5th-gen tensor cores, thread-per-row flavour — behind distegnn_edge_layer_fwd_t16
// (the production symbol distegnn_edge_layer_fwd runs the column-split flavour, edge_layer_cs.cu; this one is its twin).
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

#include <type_traits>

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
                               + T16_GROUPS * TILE_M * 16 // next tile's row | col | edge_attr[0..1] (T16_IDX_STAGE)
                               + 128;                     // mbarriers + tmem base
constexpr uint32_t T16_LBO = 1024, T16_SBO = 128;         // fp16 K-major no-swizzle: 8 rows x 8 halfs per core matrix
constexpr float T16_RANGE = 3.0e4f;
#ifndef T16_CHUNK_UNROLL
#define T16_CHUNK_UNROLL 2     // chunks (of 16 columns) unrolled per stage loop: trades code size (I-cache) for ILP
#endif
constexpr int kChunkUnroll = T16_CHUNK_UNROLL;
// A/B knobs (python -m distegnn_b200.build --variant TAG --defs=...): the defaults are the measured winners.
#ifndef T16_Q_COPY
#define T16_Q_COPY 0        // neighbour rows Q[col] -> shared: 0 = one TMA bulk copy per edge (UBLKCP needs uniform
#endif                      // operands: ptxas emits a 9-instruction loop over the 32 lanes), 1 = 16 LDGSTS per thread
                            // (measured: 3.72 ms vs 2.89 ms for the TMA form — scattered 16-byte LDGSTS lose)
#ifndef T16_SEGSUM
#define T16_SEGSUM 1        // segment sum of m: 0 = thread per (column, half tile), scalar; 1 = thread per (column
#endif                      // pair, quarter tile): LDS.64 + FADD2, one RED.v2 per run
#ifndef T16_IDX_STAGE
#define T16_IDX_STAGE 1     // next tile's (row, col, edge_attr): 0 = LDG into registers at mid-iteration (dependent
#endif                      // x4 / P-prefetch addresses stall on it), 1 = LDGSTS into shared one stage earlier
constexpr std::false_type kFast{};   // silu4p flavour tags (common.cuh)
constexpr std::true_type kSafe{};

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

// 16 fp32 values held as 8 register pairs (·s) -> 8 packed hi words + 8 packed lo words; `mx` tracks the running
// max of the hi halves.  Per pair: F2FP, 2 HADD2.F32 (unpack), FADD2, F2FP, HMNMX2.
template <bool SCALED>
__device__ __forceinline__ void split16(const f32x2 (&v)[8], float s, uint32_t (&hi)[8], uint32_t (&lo)[8], __half2& mx) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const f32x2 x = SCALED ? mul2(v[j], bc2(s)) : v[j];
        float x0, x1, l0, l1;
        upk2(x, x0, x1);
        const __half2 h = __floats2half2_rn(x0, x1);            // x0 -> low half (even k)
        const float2 hf = __half22float2(h);
        upk2(sub2(x, pk2(hf.x, hf.y)), l0, l1);
        const __half2 l = __floats2half2_rn(l0, l1);
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
    int* nidx_all = reinterpret_cast<int*>(rmask_all + T16_GROUPS * 4);               // [4][row 128 | col 128 | ea 256]
    uint64_t* bars = reinterpret_cast<uint64_t*>(nidx_all + T16_GROUPS * TILE_M * 4);  // [4][2]
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
    for (int i = tid; i < DISTEGNN_MAX_EDGE_ATTR * H; i += T16_THREADS) w1es[i] = i < A * H ? a.w1e[i] : 0.f;   // zero rows: the generic (AT < 0) loop runs over all AMAX
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
    int* nrow_s = nidx_all + grp * TILE_M * 4;
    int* ncol_s = nrow_s + TILE_M;
    float* nea_s = reinterpret_cast<float*>(ncol_s + TILE_M);      // [128][2]
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
    constexpr bool kIdxStage = T16_IDX_STAGE && (AT == 1 || AT == 2 || AT == 0);
    auto prefetch_q = [&](int64_t tl, int r, int c) {
#if T16_Q_COPY == 0
        if (tl < num_tiles) {
            if (t == 0) {
                const int64_t nvalid = min((int64_t)TILE_M, a.E - tl * TILE_M);
                mbar_expect_tx(qbar, (uint32_t)nvalid * (H * 4));
            }
            if (r >= 0) bulk_g2s(myq, a.Q + (size_t)c * H, H * 4, qbar);
        }
#else
        if (tl < num_tiles && r >= 0) {              // thread t copies (and later reads) row t: no barrier needed
            const float* src = a.Q + (size_t)c * H;
#pragma unroll
            for (int k = 0; k < H / 4; ++k) cp_async16(myq + 4 * k, src + 4 * k);
        }
        cp_async_commit();
#endif
    };
    // T16_IDX_STAGE: start the copy of tile tl's (row, col, edge_attr) of this thread's edge into shared memory
    auto stage_idx = [&](int64_t tl) {
        const int64_t e = tl * TILE_M + t;
        if (tl < num_tiles && e < a.E) {
            cp_async4(nrow_s + t, a.row + e);
            cp_async4(ncol_s + t, a.col + e);
            if (AT == 1) cp_async4(nea_s + 2 * t, a.ea + e);
            if (AT == 2) cp_async8(nea_s + 2 * t, a.ea + e * 2);
        }
        cp_async_commit();
    };
    auto take_idx = [&](int64_t tl, int& r, int& rr, int& c, float (&ea)[AMAX]) {
        const int64_t e = tl * TILE_M + t;
        r = -1;
        rr = 0;
        c = 0;
#pragma unroll
        for (int k = 0; k < AMAX; ++k) ea[k] = 0.f;
        cp_async_wait_all();
        if (tl < num_tiles && e < a.E) {
            r = nrow_s[t];
            c = ncol_s[t];
            rr = r;
            if (AT == 1) ea[0] = nea_s[2 * t];
            if (AT == 2) {
                const float2 v = *reinterpret_cast<const float2*>(nea_s + 2 * t);
                ea[0] = v.x;
                ea[AMAX - 1] = v.y;
            }
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
#if T16_Q_COPY == 0
        mbar_wait(qbar, (uint32_t)(it & 1));
        __syncwarp();
#else
        cp_async_wait_all();                   // this thread's own row of Q is in shared memory
#endif
        if (kIdxStage) stage_idx(tile + stride);
        const float* prow = a.P + (size_t)rr_c * H;
        const f32x2 rad2 = bc2(radial);
        float qmax = 0.f;                  // SiLU batch-reciprocal range guard (common.cuh silu4p)
        auto pre_chunk = [&](int c, f32x2 (&v)[8], auto safe) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const int cc = 16 * c + 4 * j4;
                const ulonglong2 pp = __ldg(reinterpret_cast<const ulonglong2*>(prow + cc));
                const ulonglong2 qq = *reinterpret_cast<const ulonglong2*>(myq + cc);
                const ulonglong2 wr = *reinterpret_cast<const ulonglong2*>(w1rs + cc);
                f32x2 p0 = fma2(rad2, wr.x, add2(pp.x, qq.x)), p1 = fma2(rad2, wr.y, add2(pp.y, qq.y));
#pragma unroll
                for (int k = 0; k < AMAX; ++k)
                    if (AT < 0 || k < A) {         // AT < 0: ea_c[k] = 0 and zero weight rows beyond A (a predicated
                                                   // FFMA2 chain here crashes ptxas 12.9 at -O2 and above)
                        const ulonglong2 we = *reinterpret_cast<const ulonglong2*>(w1es + k * H + cc);
                        const f32x2 e2 = bc2(ea_c[k]);
                        p0 = fma2(e2, we.x, p0);
                        p1 = fma2(e2, we.y, p1);
                    }
                silu4p<decltype(safe)::value>(p0, p1, qmax);
                v[2 * j4] = p0;
                v[2 * j4 + 1] = p1;
            }
        };
        float inv_s1 = 1.0f;
        {
            __half2 mx = __floats2half2_rn(0.f, 0.f);
#pragma unroll kChunkUnroll
            for (int c = 0; c < 4; ++c) {
                f32x2 v[8];
                uint32_t hi[8], lo[8];
                pre_chunk(c, v, kFast);
                split16<false>(v, 1.0f, hi, lo, mx);
                tmem_st8(lane_off + tA_hi + 8 * c, hi);
                tmem_st8(lane_off + tA_lo + 8 * c, lo);
            }
            if (__any_sync(FULL, row_overflow(mx) || silu_q_overflow(qmax))) {   // cold: a row of this warp leaves the
                                                                                  // fp16 range, or the SiLU batch guard fired
                float fm = 0.f, sc;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    f32x2 v[8];
                    pre_chunk(c, v, kSafe);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float v0, v1;
                        upk2(v[j], v0, v1);
                        fm = fmaxf(fm, fmaxf(v0, v1));
                    }
                }
                range_scale(fm, sc, inv_s1);
                wait_st();
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    f32x2 v[8];
                    uint32_t hi[8], lo[8];
                    pre_chunk(c, v, kSafe);
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
        if (kIdxStage) take_idx(tile + stride, row_n, rr_n, col_n, ea_n);
        else load_edge(tile + stride, row_n, rr_n, col_n, ea_n);
        prefetch_l1(a.P + (size_t)rr_n * H);
        prefetch_l1(a.P + (size_t)rr_n * H + 32);
        const float4 xi_n = ldg4(a.x4 + (size_t)rr_n * 4), xj_n = ldg4(a.x4 + (size_t)col_n * 4);

        mbar_wait(mbar, 0);
        __syncwarp();
        fence_after_sync();

        // ---- stage 2: m = SiLU(D/s + b2) -> row to shared (segment sum) and fp16 hi/lo -> TMEM ---------------
        qmax = 0.f;
        auto m_chunk = [&](int c, f32x2 (&v)[8], bool store, float inv_s, auto safe) {
            uint32_t d[16];
            tmem_ld16(lane_off + tD + 16 * c, d);
            wait_ld();
            const f32x2 is2 = bc2(inv_s);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const int cc = 16 * c + 4 * j4;
                const ulonglong2 bb = *reinterpret_cast<const ulonglong2*>(b2s + cc);
                f32x2 m0 = fma2(pk2u(d[4 * j4 + 0], d[4 * j4 + 1]), is2, bb.x);
                f32x2 m1 = fma2(pk2u(d[4 * j4 + 2], d[4 * j4 + 3]), is2, bb.y);
                silu4p<decltype(safe)::value>(m0, m1, qmax);
                if (store) *reinterpret_cast<ulonglong2*>(myq + cc) = make_ulonglong2(m0, m1);
                v[2 * j4] = m0;
                v[2 * j4 + 1] = m1;
            }
        };
        float inv_s2 = 1.0f;
        {
            __half2 mx = __floats2half2_rn(0.f, 0.f);
#pragma unroll kChunkUnroll
            for (int c = 0; c < 4; ++c) {
                f32x2 v[8];
                uint32_t hi[8], lo[8];
                m_chunk(c, v, need_m, inv_s1, kFast);
                split16<false>(v, 1.0f, hi, lo, mx);
                tmem_st8(lane_off + tA_hi + 8 * c, hi);
                tmem_st8(lane_off + tA_lo + 8 * c, lo);
            }
            if (__any_sync(FULL, row_overflow(mx) || silu_q_overflow(qmax))) {      // cold
                float fm = 0.f, sc;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    f32x2 v[8];
                    m_chunk(c, v, need_m, inv_s1, kSafe);      // also rewrites the m row in shared memory
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float v0, v1;
                        upk2(v[j], v0, v1);
                        fm = fmaxf(fm, fmaxf(v0, v1));
                    }
                }
                range_scale(fm, sc, inv_s2);
                wait_st();
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    f32x2 v[8];
                    uint32_t hi[8], lo[8];
                    m_chunk(c, v, false, inv_s1, kSafe);
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
#if T16_SEGSUM == 1
            // warp wq <-> edges 32wq .. 32wq+31 of the tile, lane <-> columns 2·lane, 2·lane+1: per edge one LDS.64 and
            // one FADD2; the run structure (bit mask of run starts) is warp-uniform, one RED.v2 per run and lane
            const float* colp = qb + (32 * wq) * T16_QROW + 2 * lane;
            uint32_t M = rmask[wq] | 1u;
            while (M) {
                const int e0 = __ffs((int)M) - 1;
                M &= M - 1;
                const int e1 = M ? __ffs((int)M) - 1 : 32;
                f32x2 s0 = 0ull, s1 = 0ull, s2 = 0ull, s3 = 0ull;          // (+0, +0)
                int e = e0;
                for (; e + 3 < e1; e += 4) {
                    s0 = add2(s0, *reinterpret_cast<const f32x2*>(colp + e * T16_QROW));
                    s1 = add2(s1, *reinterpret_cast<const f32x2*>(colp + (e + 1) * T16_QROW));
                    s2 = add2(s2, *reinterpret_cast<const f32x2*>(colp + (e + 2) * T16_QROW));
                    s3 = add2(s3, *reinterpret_cast<const f32x2*>(colp + (e + 3) * T16_QROW));
                }
                for (; e < e1; ++e) s0 = add2(s0, *reinterpret_cast<const f32x2*>(colp + e * T16_QROW));
                const int r = srow[32 * wq + e0];
                if (r >= 0) {
                    float v0, v1;
                    upk2(add2(add2(s0, s1), add2(s2, s3)), v0, v1);
                    red_add_v2(a.agg_m + (size_t)r * H + 2 * lane, v0, v1);
                }
            }
#else
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
#endif
#if T16_Q_COPY == 0
            fence_proxy_async_smem();          // generic accesses to qb ordered before the TMA refill below
#endif
        }
        named_bar(bar_id, T16_GROUP);          // whole group done with the staging buffer
        prefetch_q(tile + stride, row_n, col_n);

        mbar_wait(mbar, 1);
        __syncwarp();
        fence_after_sync();

        // ---- stage 3: φ = w3·SiLU(D/s + bc); Δx·φ summed per destination row --------------------------------
        f32x2 ph01, ph23;                                       // four independent FMA chains in two register pairs
        qmax = 0.f;
        auto phi_pass = [&](auto safe) {
            ph01 = bc2(0.f);
            ph23 = bc2(0.f);
            const f32x2 is2 = bc2(inv_s2);
#pragma unroll kChunkUnroll
            for (int c = 0; c < 4; ++c) {
                uint32_t d[16];
                tmem_ld16(lane_off + tD + 16 * c, d);
                wait_ld();
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const int cc = 16 * c + 4 * j4;
                    const ulonglong2 bb = *reinterpret_cast<const ulonglong2*>(bcs + cc);
                    const ulonglong2 ww = *reinterpret_cast<const ulonglong2*>(w3s + cc);
                    f32x2 s0 = fma2(pk2u(d[4 * j4 + 0], d[4 * j4 + 1]), is2, bb.x);
                    f32x2 s1 = fma2(pk2u(d[4 * j4 + 2], d[4 * j4 + 3]), is2, bb.y);
                    silu4p<decltype(safe)::value>(s0, s1, qmax);
                    ph01 = fma2(s0, ww.x, ph01);
                    ph23 = fma2(s1, ww.y, ph23);
                }
            }
        };
        phi_pass(kFast);
        if (kSiluGuard && __any_sync(FULL, silu_q_overflow(qmax))) phi_pass(kSafe);      // cold
        float ph0, ph1, ph2, ph3;
        upk2(ph01, ph0, ph1);
        upk2(ph23, ph2, ph3);
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

extern "C" int distegnn_edge_layer_fwd_t16(int64_t n_nodes, int64_t n_edges, int A, int C, int Na, unsigned flags,
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
