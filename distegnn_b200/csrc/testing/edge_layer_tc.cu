// Real<->real edge stage on the 5th-gen tensor cores (tcgen05 + TMEM), the production edge kernel.
// Same math and outputs as edge_layer.cu (the fp32-FMA version kept as distegnn_edge_layer_fwd_simt for
// cross-checks); replaces reference models/FastEGNN.py:237-246,144-150,169-177,206 and the scatter_add_ of
// :322-337.
//
// One CTA per SM, 512 threads = 2 independent tile groups of 8 warps.  A group owns one 128-edge tile at a
// time; edge r of the tile lives in TMEM lane r and is handled by TWO threads (warps q and q+4 of the group
// share TMEM lane quarter q and split the 64 feature columns 32/32 — twice the warps for the same TMEM):
//   stage 0  the 256-byte neighbour rows Q[col] of the NEXT tile are fetched by the TMA engine
//            (cp.async.bulk global->shared, one copy per edge, completion on an mbarrier) while the
//            current tile computes — the [E,64] gathers of the reference never exist;
//   stage 1  a1 = SiLU(P[row] + Q[col] + w_r·r + W_e·a)  -> split hi/lo (3xTF32) -> tcgen05.st to TMEM;
//   MMA 1    D = a1·W2ᵀ : 24 tcgen05.mma (M128,N64,K8, kind::tf32; lo·hi + hi·lo + hi·hi) issued by one
//            thread, B = W2 resident in shared memory (no-swizzle K-major descriptor), D in TMEM;
//   stage 2  m = SiLU(D + b2) (tcgen05.ld, row per thread) -> m row to shared (for the segment sum) and
//            hi/lo back to TMEM;
//   MMA 2    D = m·Wcᵀ ; while it runs the WG segment-sums m over runs of equal destination row
//            (column walk over the shared tile, RED.ADD per (run, column));
//   stage 3  φ = w3·SiLU(D + bc), Δx·φ reduced over runs of equal row with warp shuffles, RED.ADD to agg_x.
// While one WG waits on its MMAs the other WG's SIMT stages use the issue slots, and vice versa.
#include "common.cuh"
#include "umma.cuh"

namespace degnn {

struct EdgeTcArgs {
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

constexpr int TC_THREADS = 512;                           // 2 tile groups x 8 warps
constexpr int GROUP_THREADS = 256;                        // warps (q, half): TMEM lane quarter q, column half
constexpr int QROW = 68;                                  // floats per staged row (272 B: 16B-aligned, bank-shifted)
constexpr int QBUF_FLOATS = TILE_M * QROW;                // 8704
constexpr int KEXT = 72;                                  // K of the tile GEMMs: 64 features + a ones column (bias) + pad
constexpr int BEXT_FLOATS = (KEXT / 4) * 256;             // 18 K-chunks x 1 KB = 4608 floats per weight matrix
constexpr int TC_SMEM_BYTES = 4 * BEXT_FLOATS * 4         // [W2|b2] hi/lo, [Wc|bc] hi/lo (UMMA canonical layout)
                              + 2 * 2 * QBUF_FLOATS * 4   // 2 groups x 2 staging buffers
                              + (2 * H + DISTEGNN_MAX_EDGE_ATTR * H) * 4   // w3, w1r, w1e
                              + 2 * 2 * TILE_M * 4        // srow per group, double-buffered by tile parity
                              + 2 * TILE_M * 4            // partial φ per group
                              + 64;                       // mbarriers + tmem base
// TMEM columns of one tile group (256 per group): A_hi [0,72)  A_lo [72,144)  D [144,208)
constexpr uint32_t TCOL_AHI = 0, TCOL_ALO = 72, TCOL_D = 144;

// B operand [n][k], k < 72: W[n][k] for k < 64, bias[n] at k = 64, zeros beyond (A carries 1.0 in column 64)
__device__ __forceinline__ void stage_weight_umma(float* hi, float* lo, const float* __restrict__ wt_kmajor,
                                                  const float* __restrict__ bias, int tid, int nthreads) {
    for (int i = tid; i < KEXT * H; i += nthreads) {
        const int k = i >> 6, n = i & 63;
        const float v = (k < H) ? __ldg(wt_kmajor + i) : (k == H ? __ldg(bias + n) : 0.f);
        uint32_t h, l;
        umma::split_tf32(v, h, l);
        const uint32_t o = umma::b_elem_offset(n, k);
        hi[o] = __uint_as_float(h);
        lo[o] = __uint_as_float(l);
    }
}

// 27 MMAs: D = Alo·Bhiᵀ + Ahi·Bloᵀ + Ahi·Bhiᵀ over K = 72 (small terms first), then commit to `bar`
__device__ __forceinline__ void issue_gemm_3xtf32(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi,
                                                  uint64_t b_lo, uint32_t idesc, uint64_t* bar) {
    constexpr uint64_t KSTEP = (2 * umma::B_LBO) >> 4;   // descriptor start-address units per K=8 step
    constexpr int NK = KEXT / 8;
#pragma unroll
    for (int ks = 0; ks < NK; ++ks) umma::mma_tf32_ts(d, a_lo + 8 * ks, b_hi + ks * KSTEP, idesc, ks > 0);
#pragma unroll
    for (int ks = 0; ks < NK; ++ks) umma::mma_tf32_ts(d, a_hi + 8 * ks, b_lo + ks * KSTEP, idesc, 1);
#pragma unroll
    for (int ks = 0; ks < NK; ++ks) umma::mma_tf32_ts(d, a_hi + 8 * ks, b_hi + ks * KSTEP, idesc, 1);
    umma::mma_commit(bar);
}

template <int AT>   // AT = edge_attr_nf when 0..2, else -1 (runtime loop)
__global__ void __launch_bounds__(TC_THREADS, 1) edge_layer_tc_kernel(const EdgeTcArgs a) {
    using namespace umma;
    constexpr int AMAX = AT >= 0 ? (AT > 0 ? AT : 1) : DISTEGNN_MAX_EDGE_ATTR;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    float* W2hi = reinterpret_cast<float*>(smem_raw);
    float* W2lo = W2hi + BEXT_FLOATS;
    float* Wchi = W2lo + BEXT_FLOATS;
    float* Wclo = Wchi + BEXT_FLOATS;
    float* qbufs = Wclo + BEXT_FLOATS;                        // [2 groups][2][QBUF_FLOATS]
    float* w3s = qbufs + 4 * QBUF_FLOATS;
    float* w1rs = w3s + H;
    float* w1es = w1rs + H;                                   // [A][64]
    int* srow_all = reinterpret_cast<int*>(w1es + DISTEGNN_MAX_EDGE_ATTR * H);   // [2 groups][2][128]
    float* phi_all = reinterpret_cast<float*>(srow_all + 4 * TILE_M);             // [2 groups][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(phi_all + 2 * TILE_M);          // [2 groups][3]
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 6);

    const int tid = threadIdx.x;
    const int grp = tid >> 8;              // tile group 0/1 (each owns one 128-edge tile at a time)
    const int u = tid & 255;               // thread index inside the group
    const int lane = tid & 31;
    const int w8 = u >> 5;                 // warp inside the group
    const int wq = w8 & 3;                 // TMEM lane quarter (== warp id % 4, a hardware rule)
    const int half = w8 >> 2;              // which 32 of the 64 feature columns this warp handles
    const int t = 32 * wq + lane;          // edge (row) of the tile owned by this thread
    const int cb = 32 * half;              // first feature column of this thread
    const int A = AT >= 0 ? AT : a.A;
    const bool normalize = a.flags & DISTEGNN_FLAG_NORMALIZE;
    const bool need_m = !(a.flags & DISTEGNN_FLAG_LAST);

    // ---- one-time setup -------------------------------------------------------------------------
    stage_weight_umma(W2hi, W2lo, a.w2, a.b2, tid, TC_THREADS);
    stage_weight_umma(Wchi, Wclo, a.wc, a.bc, tid, TC_THREADS);
    if (tid < H) {
        w3s[tid] = a.w3[tid];
        w1rs[tid] = a.w1r[tid];
    }
    for (int i = tid; i < A * H; i += TC_THREADS) w1es[i] = a.w1e[i];
    if (tid == 0) {
        for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    __syncwarp();
    if ((tid >> 5) == 0) tmem_alloc(tmem_base_s, 512);
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();

    const uint32_t tbase = *tmem_base_s;
    const uint32_t col0 = tbase + (uint32_t)grp * 256u;            // this group's TMEM columns
    const uint32_t tA_hi = col0 + TCOL_AHI, tA_lo = col0 + TCOL_ALO, tD = col0 + TCOL_D;
    const uint32_t lane_off = ((uint32_t)(32 * wq)) << 16;          // this warp's TMEM lane quarter
    const uint32_t idesc = make_idesc_tf32(128, 64);
    const uint64_t dW2hi = make_b_desc(smem_u32(W2hi), B_LBO, B_SBO), dW2lo = make_b_desc(smem_u32(W2lo), B_LBO, B_SBO);
    const uint64_t dWchi = make_b_desc(smem_u32(Wchi), B_LBO, B_SBO), dWclo = make_b_desc(smem_u32(Wclo), B_LBO, B_SBO);
    float* qbuf0 = qbufs + (grp * 2 + 0) * QBUF_FLOATS;
    int* srow2 = srow_all + grp * 2 * TILE_M;
    float* phis = phi_all + grp * TILE_M;
    uint64_t* qbar = bars + grp * 3;       // [2]
    uint64_t* mbar = bars + grp * 3 + 2;
    const uint32_t bar_id = 1 + grp;

    // the bias column of the A operand: A[:,64] = 1, A[:,65:72] = 0 — written once, never touched again
    if (half == 0) {
        uint32_t one[8] = {0x3f800000u, 0u, 0u, 0u, 0u, 0u, 0u, 0u}, zero[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        tmem_st8(lane_off + tA_hi + H, one);
        tmem_st8(lane_off + tA_lo + H, zero);
        wait_st();
    }

    const int64_t num_tiles = (a.E + TILE_M - 1) / TILE_M;
    const int64_t stride = (int64_t)gridDim.x * 2;
    int64_t tile = (int64_t)blockIdx.x * 2 + grp;

    // per-edge state of the tile about to be processed; invalid (tail) edges carry row -1 but read node 0
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
    auto geometry = [&](int rr, int c, float& gx, float& gy, float& gz, float& rad) {
        const float4 xi = ldg4(a.x4 + (size_t)rr * 4), xj = ldg4(a.x4 + (size_t)c * 4);
        gx = xi.x - xj.x; gy = xi.y - xj.y; gz = xi.z - xj.z;
        rad = gx * gx + gy * gy + gz * gz;
        if (normalize) {
            const float inv = 1.0f / (sqrtf(rad) + 1e-8f);
            gx *= inv; gy *= inv; gz *= inv;
        }
    };
    // neighbour rows Q[col] of tile `tl` -> staging buffer, one TMA bulk copy per edge (column-half-0 warps)
    auto prefetch_q = [&](int64_t tl, int r, int c, float* dst, uint64_t* bar) {
        if (tl < num_tiles && half == 0) {
            if (t == 0) {
                const int64_t nvalid = min((int64_t)TILE_M, a.E - tl * TILE_M);
                mbar_expect_tx(bar, (uint32_t)nvalid * (H * 4));
            }
            if (r >= 0) bulk_g2s(dst + t * QROW, a.Q + (size_t)c * H, H * 4, bar);
        }
    };

    if (tile < num_tiles) {
        int col_c;
        load_edge(tile, row_c, rr_c, col_c, ea_c);
        prefetch_q(tile, row_c, col_c, qbuf0, &qbar[0]);
        geometry(rr_c, col_c, dx, dy, dz, radial);
    }

    for (int it = 0; tile < num_tiles; ++it, tile += stride) {
        const int b = it & 1;
        float* qb = qbuf0 + b * QBUF_FLOATS;
        int* srow = srow2 + b * TILE_M;    // read by this tile's segment sum until the next-but-one tile

        // ---- stage 1: a1 = SiLU(P_i + Q_j + w_r·r + W_e·a) for this thread's 32 columns ---------------
        mbar_wait(&qbar[b], (uint32_t)((it >> 1) & 1));
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int cc = cb + 16 * c;
            uint32_t hi[16], lo[16];
            const float* prow = a.P + (size_t)rr_c * H + cc;
            const float* qrow = qb + t * QROW + cc;
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 p = ldg4(prow + 4 * j4);
                const float4 q = *reinterpret_cast<const float4*>(qrow + 4 * j4);
                float4 pre = fma4(radial, *reinterpret_cast<const float4*>(w1rs + cc + 4 * j4), add4(p, q));
#pragma unroll
                for (int k = 0; k < AMAX; ++k)
                    if (k < A) pre = fma4(ea_c[k], *reinterpret_cast<const float4*>(w1es + k * H + cc + 4 * j4), pre);
                pre = silu4(pre);
                split_tf32(pre.x, hi[4 * j4 + 0], lo[4 * j4 + 0]);
                split_tf32(pre.y, hi[4 * j4 + 1], lo[4 * j4 + 1]);
                split_tf32(pre.z, hi[4 * j4 + 2], lo[4 * j4 + 2]);
                split_tf32(pre.w, hi[4 * j4 + 3], lo[4 * j4 + 3]);
            }
            tmem_st16(lane_off + tA_hi + cc, hi);
            tmem_st16(lane_off + tA_lo + cc, lo);
        }
        wait_st();
        if (half == 0) srow[t] = row_c;
        fence_before_sync();
        named_bar(bar_id, GROUP_THREADS);   // A complete; group is done with the other staging buffer and with D

        // ---- MMA 1; meanwhile fetch the next tile's edges, neighbour rows (TMA) and geometry -------------
        if (u == 0) {
            fence_after_sync();
            issue_gemm_3xtf32(tD, tA_hi, tA_lo, dW2hi, dW2lo, idesc, mbar);
        }
        __syncwarp();
        int row_n, rr_n, col_n;
        float ea_n[AMAX];
        load_edge(tile + stride, row_n, rr_n, col_n, ea_n);
        prefetch_q(tile + stride, row_n, col_n, qbuf0 + (b ^ 1) * QBUF_FLOATS, &qbar[b ^ 1]);
        prefetch_l1(a.P + (size_t)rr_n * H + cb);             // this thread's 128-byte half of P[row]
        const float4 xi_n = ldg4(a.x4 + (size_t)rr_n * 4), xj_n = ldg4(a.x4 + (size_t)col_n * 4);

        mbar_wait(mbar, 0);
        __syncwarp();
        fence_after_sync();

        // ---- stage 2: m = SiLU(D)   (b2 rides in the GEMM through the ones column) ---------------------
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int cc = cb + 16 * c;
            uint32_t d[16], hi[16], lo[16];
            tmem_ld16(lane_off + tD + cc, d);
            wait_ld();
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                float4 m;
                m.x = silu(__uint_as_float(d[4 * j4 + 0]));
                m.y = silu(__uint_as_float(d[4 * j4 + 1]));
                m.z = silu(__uint_as_float(d[4 * j4 + 2]));
                m.w = silu(__uint_as_float(d[4 * j4 + 3]));
                if (need_m) *reinterpret_cast<float4*>(qb + t * QROW + cc + 4 * j4) = m;
                split_tf32(m.x, hi[4 * j4 + 0], lo[4 * j4 + 0]);
                split_tf32(m.y, hi[4 * j4 + 1], lo[4 * j4 + 1]);
                split_tf32(m.z, hi[4 * j4 + 2], lo[4 * j4 + 2]);
                split_tf32(m.w, hi[4 * j4 + 3], lo[4 * j4 + 3]);
            }
            tmem_st16(lane_off + tA_hi + cc, hi);
            tmem_st16(lane_off + tA_lo + cc, lo);
        }
        wait_st();
        fence_before_sync();
        named_bar(bar_id, GROUP_THREADS);   // m tile visible in shared, A operand complete, D fully read

        // ---- MMA 2 (φ head) overlapped with the segment sum of m ---------------------------------
        if (u == 0) {
            fence_after_sync();
            issue_gemm_3xtf32(tD, tA_hi, tA_lo, dWchi, dWclo, idesc, mbar);
        }
        __syncwarp();
        if (need_m) {
            // thread (column c, quarter of the tile): runs of equal destination row -> one RED per run
            const int c = u & 63, eb = (u >> 6) * 32;
            const float* col = qb + eb * QROW + c;
            int cur = srow[eb];
            float s = 0.f;
#pragma unroll 8
            for (int e = 0; e < 32; ++e) {
                const int r = srow[eb + e];
                if (r != cur) {
                    if (cur >= 0) atomicAdd(a.agg_m + (size_t)cur * H + c, s);
                    s = 0.f;
                    cur = r;
                }
                s += col[e * QROW];
            }
            if (cur >= 0) atomicAdd(a.agg_m + (size_t)cur * H + c, s);
            fence_proxy_async_smem();   // generic accesses to qb ordered before its next TMA refill
        }

        mbar_wait(mbar, 1);
        __syncwarp();
        fence_after_sync();

        // ---- stage 3: φ = w3·SiLU(D); Δx·φ summed per destination row ---------------------------------
        float phi = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int cc = cb + 16 * c;
            uint32_t d[16];
            tmem_ld16(lane_off + tD + cc, d);
            wait_ld();
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 ww = *reinterpret_cast<const float4*>(w3s + cc + 4 * j4);
                phi = fmaf(silu(__uint_as_float(d[4 * j4 + 0])), ww.x, phi);
                phi = fmaf(silu(__uint_as_float(d[4 * j4 + 1])), ww.y, phi);
                phi = fmaf(silu(__uint_as_float(d[4 * j4 + 2])), ww.z, phi);
                phi = fmaf(silu(__uint_as_float(d[4 * j4 + 3])), ww.w, phi);
            }
        }
        fence_before_sync();   // D reads ordered before the next tile's MMA (after the next named barrier)
        if (half == 1) phis[t] = phi;
        named_bar(bar_id, GROUP_THREADS);
        if (half == 0) {
            phi += phis[t];
            float sx = dx * phi, sy = dy * phi, sz = dz * phi;
            // inclusive segmented scan over the warp (keys sorted): lane adds lanes below with the same row
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

        // ---- roll the prefetched edge into place ------------------------------------------------------
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

extern "C" int distegnn_edge_layer_fwd_tf32(int64_t n_nodes, int64_t n_edges, int A, int C, int Na, unsigned flags,
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
    EdgeTcArgs a;
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
    int64_t grid = (tiles + 1) / 2;
    if (grid > sm_count()) grid = sm_count();
    auto launch = [&](auto kern) {
        ensure_dynamic_smem((const void*)kern, (int)TC_SMEM_BYTES);
        kern<<<(unsigned)grid, TC_THREADS, TC_SMEM_BYTES, (cudaStream_t)stream>>>(a);
    };
    switch (A) {
        case 0: launch(edge_layer_tc_kernel<0>); break;
        case 1: launch(edge_layer_tc_kernel<1>); break;
        case 2: launch(edge_layer_tc_kernel<2>); break;
        default: launch(edge_layer_tc_kernel<-1>); break;
    }
    DEGNN_CHECK_LAUNCH();
    return DISTEGNN_OK;
}
