// Real<->virtual stage on the tensor cores with the 3xTF32 split (earlier production kernel, now
// distegnn_virtual_layer_fwd_tf32: A/B timing and a third independent implementation for cross-checks).  Same math/outputs as virtual_layer.cu (kept as ..._simt for cross-checks);
// replaces reference models/FastEGNN.py:252-253,154-163,180,191-193,207,220-223 and the global_mean_pool
// scatters at :193,:222.
//
// Rows of a tile are (node, channel) pairs: TN = 128 / C nodes per tile, row = n_local*C + c, one TMEM lane per
// row.  One CTA per SM, 512 threads = 2 tile groups x 8 warps; warps q and q+4 of a group share TMEM lane
// quarter q and split the feature columns (stage 1/2: 32+32 of 64; stage 3: one coordinate head each).
//   stage 1  a1 = SiLU(Hn[node] + G[graph,c] + w_r·‖ΔX‖)          -> hi/lo -> TMEM
//   MMA 1    D[:, 0:64]  = a1·W2vᵀ                                   (24 x tcgen05.mma M128 N64 K8 tf32)
//   stage 2  mv = SiLU(D + b2v) -> shared tile (for the pools) and hi/lo -> TMEM
//   MMA 2    D[:, 0:128] = mv·[Wxv;WX]ᵀ                              (24 x tcgen05.mma M128 N128 K8)
//            overlapped with: agg_v[node] = mean_c mv, per-graph Σ_i mv accumulated on chip
//   stage 3  φ_xv = w3xv·SiLU(D[:, :64]+bxv) (warps 0-3), φ_X = w3x·SiLU(D[:, 64:]+bx) (warps 4-7);
//            trans_v[node] = mean_c(−ΔX·φ_xv);  per-graph Σ_i ΔX·φ_X accumulated on chip.
// Per-graph sums live in shared memory per group and are flushed to `vsum` with RED.ADD when the group moves
// to another graph and at the end (tiles that straddle graphs use RED.ADD directly).
#include "common.cuh"
#include "umma.cuh"

namespace degnn {

struct VirtTcArgs {
    int64_t N;
    int B, C;
    unsigned flags;
    const int32_t* batch;
    const float* x4;
    const float* Hn;
    const float* Xv;
    const float* G;
    const float* w1r;
    const float* w2; const float* b2;
    const float* wxv; const float* bxv; const float* w3xv;
    const float* wx; const float* bx; const float* w3x;
    float* agg_v;
    float* trans_v;
    float* vsum;
};

constexpr int VT_THREADS = 512, VT_GROUP = 256;
constexpr int VT_ROW = 68;
constexpr int VT_MAXC = DISTEGNN_MAX_CHANNELS;
constexpr uint32_t VT_LBO2 = 2048;                         // K-chunk stride of the 128-row head matrix
constexpr int VT_SMEM_FLOATS = 2 * 4096                    // W2v hi/lo
                               + 2 * 8192                  // [Wxv;WX] hi/lo
                               + 2 * TILE_M * VT_ROW       // mv tile per group
                               + 2 * VT_MAXC * H           // Σ mv accumulators per group
                               + 2 * 4 * VT_MAXC           // Σ ΔX·φ_X accumulators per group
                               + 6 * H                     // w1r, b2v, [bxv|bx], [w3xv|w3x]
                               + 2 * TILE_M * 4            // ΔX per row per group
                               + 2 * 2 * TILE_M            // φ_xv, φ_X per row per group
                               + 2 * TILE_M                // graph id per local node per group (int)
                               + 16;                       // mbarriers + tmem base
constexpr int VT_SMEM_BYTES = VT_SMEM_FLOATS * 4;

__device__ __forceinline__ void stage_w_generic(float* hi, float* lo, const float* __restrict__ wt_kmajor,
                                                int n_off, uint32_t lbo_floats, int tid, int nthreads) {
    // wt_kmajor[k*64+n] = W[n][k] -> B element (n_off+n, k) at (k/4)*lbo + ((n_off+n)/8)*32 + ((n_off+n)%8)*4 + k%4
    for (int i = tid; i < H * H; i += nthreads) {
        const int k = i >> 6, n = (i & 63) + n_off;
        uint32_t h, l;
        umma::split_tf32(__ldg(wt_kmajor + i), h, l);
        const uint32_t o = (uint32_t)(k >> 2) * lbo_floats + (uint32_t)(n >> 3) * 32u + (uint32_t)(n & 7) * 4u + (k & 3);
        hi[o] = __uint_as_float(h);
        lo[o] = __uint_as_float(l);
    }
}

template <int NK, uint64_t KSTEP>
__device__ __forceinline__ void issue_3xtf32(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                             uint32_t idesc, uint64_t* bar) {
#pragma unroll
    for (int ks = 0; ks < NK; ++ks) umma::mma_tf32_ts(d, a_lo + 8 * ks, b_hi + ks * KSTEP, idesc, ks > 0);
#pragma unroll
    for (int ks = 0; ks < NK; ++ks) umma::mma_tf32_ts(d, a_hi + 8 * ks, b_lo + ks * KSTEP, idesc, 1);
#pragma unroll
    for (int ks = 0; ks < NK; ++ks) umma::mma_tf32_ts(d, a_hi + 8 * ks, b_hi + ks * KSTEP, idesc, 1);
    umma::mma_commit(bar);
}

__global__ void __launch_bounds__(VT_THREADS, 1) virtual_layer_tc_kernel(const VirtTcArgs a) {
    using namespace umma;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    float* W2hi = reinterpret_cast<float*>(smem_raw);
    float* W2lo = W2hi + 4096;
    float* WHhi = W2lo + 4096;            // heads: rows 0-63 = Wxv, 64-127 = WX
    float* WHlo = WHhi + 8192;
    float* tiles = WHlo + 8192;           // [2][128*68]
    float* accH_all = tiles + 2 * TILE_M * VT_ROW;     // [2][C][64]
    float* accX_all = accH_all + 2 * VT_MAXC * H;      // [2][3][C] (stride 4*MAXC)
    float* w1rs = accX_all + 2 * 4 * VT_MAXC;
    float* b2s = w1rs + H;
    float* bhs = b2s + H;                 // [128] bxv | bx
    float* w3s = bhs + 2 * H;             // [128] w3xv | w3x
    float* dX_all = w3s + 2 * H;          // [2][128][4]
    float* phi_all = dX_all + 2 * TILE_M * 4;          // [2][2][128]
    int* sgraph_all = reinterpret_cast<int*>(phi_all + 2 * 2 * TILE_M);   // [2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sgraph_all + 2 * TILE_M);   // [2]
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 2);

    const int tid = threadIdx.x;
    const int grp = tid >> 8, u = tid & 255, lane = tid & 31;
    const int w8 = u >> 5, wq = w8 & 3, half = w8 >> 2;
    const int t = 32 * wq + lane;          // row of the tile owned by this thread
    const int cb = 32 * half;
    const int C = a.C;
    const int K = 4 + 3 * C + H * C;
    const bool need_feat = !(a.flags & DISTEGNN_FLAG_LAST);
    const int TN = TILE_M / C;

    // ---- one-time setup ---------------------------------------------------------------------------
    stage_w_generic(W2hi, W2lo, a.w2, 0, 256, tid, VT_THREADS);
    stage_w_generic(WHhi, WHlo, a.wxv, 0, VT_LBO2 / 4, tid, VT_THREADS);
    stage_w_generic(WHhi, WHlo, a.wx, 64, VT_LBO2 / 4, tid, VT_THREADS);
    if (tid < H) {
        w1rs[tid] = a.w1r[tid];
        b2s[tid] = a.b2[tid];
        bhs[tid] = a.bxv[tid];
        bhs[H + tid] = a.bx[tid];
        w3s[tid] = a.w3xv[tid];
        w3s[H + tid] = a.w3x[tid];
    }
    for (int i = tid; i < 2 * VT_MAXC * H + 2 * 4 * VT_MAXC; i += VT_THREADS) accH_all[i] = 0.f;   // accH + accX
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    __syncwarp();
    if ((tid >> 5) == 0) tmem_alloc(tmem_base_s, 512);
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();

    const uint32_t tbase = *tmem_base_s;
    const uint32_t col0 = tbase + (uint32_t)grp * 256u;
    const uint32_t tA_hi = col0, tA_lo = col0 + 64, tD = col0 + 128;
    const uint32_t lane_off = ((uint32_t)(32 * wq)) << 16;
    const uint32_t idesc64 = make_idesc_tf32(128, 64), idesc128 = make_idesc_tf32(128, 128);
    const uint64_t dW2hi = make_b_desc(smem_u32(W2hi), B_LBO, B_SBO), dW2lo = make_b_desc(smem_u32(W2lo), B_LBO, B_SBO);
    const uint64_t dWHhi = make_b_desc(smem_u32(WHhi), VT_LBO2, B_SBO), dWHlo = make_b_desc(smem_u32(WHlo), VT_LBO2, B_SBO);
    float* tile_s = tiles + grp * TILE_M * VT_ROW;
    float* accH = accH_all + grp * VT_MAXC * H;
    float* accX = accX_all + grp * 4 * VT_MAXC;
    float* dXs = dX_all + grp * TILE_M * 4;
    float* phis = phi_all + grp * 2 * TILE_M;      // [head][row]
    int* sgraph = sgraph_all + grp * TILE_M;
    uint64_t* mbar = bars + grp;
    const uint32_t bar_id = 1 + grp;
    int cur_graph = -1;                            // graph whose sums sit in accH/accX (uniform in the group)

    auto flush = [&](int g) {                      // all threads of the group
        if (g >= 0) {
            float* dst = a.vsum + (size_t)g * K;
            if (need_feat)
                for (int i = u; i < C * H; i += VT_GROUP) {
                    atomicAdd(dst + 4 + 3 * C + i, accH[i]);
                    accH[i] = 0.f;
                }
            if (u < 3 * C) {
                atomicAdd(dst + 4 + u, accX[u]);
                accX[u] = 0.f;
            }
        }
    };

    const int64_t num_tiles = (a.N + TN - 1) / TN;
    const int64_t stride = (int64_t)gridDim.x * 2;
    for (int64_t tile = (int64_t)blockIdx.x * 2 + grp; tile < num_tiles; tile += stride) {
        const int64_t n0 = tile * TN;
        const int nvalid = (int)min((int64_t)TN, a.N - n0);
        const int rows = nvalid * C;
        if (u < TN) sgraph[u] = (u < nvalid) ? __ldg(a.batch + n0 + u) : -1;
        named_bar(bar_id, VT_GROUP);
        const int g_first = sgraph[0], g_last = sgraph[nvalid - 1];
        const bool single = (g_first == g_last);
        if (single && g_first != cur_graph) {
            flush(cur_graph);
            cur_graph = g_first;
            named_bar(bar_id, VT_GROUP);
        }

        // ---- stage 1 ------------------------------------------------------------------------------
        const bool rvalid = t < rows;
        const int nl = rvalid ? t / C : 0;
        const int ch = rvalid ? t - nl * C : 0;
        const int g = rvalid ? sgraph[nl] : g_first;
        const size_t node = (size_t)(n0 + nl);
        float vr;
        {
            const float4 xi = ldg4(a.x4 + node * 4);
            const float* Xg = a.Xv + (size_t)g * 3 * C;
            const float dx = __ldg(Xg + ch) - xi.x, dy = __ldg(Xg + C + ch) - xi.y, dz = __ldg(Xg + 2 * C + ch) - xi.z;
            vr = sqrtf(dx * dx + dy * dy + dz * dz);
            if (half == 0) *reinterpret_cast<float4*>(dXs + 4 * t) = make_float4(dx, dy, dz, 0.f);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int cc = cb + 16 * c;
            uint32_t hi[16], lo[16];
            const float* hrow = a.Hn + node * H + cc;
            const float* grow = a.G + ((size_t)g * C + ch) * H + cc;
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                float4 pre = fma4(vr, *reinterpret_cast<const float4*>(w1rs + cc + 4 * j4),
                                  add4(ldg4(hrow + 4 * j4), ldg4(grow + 4 * j4)));
                pre = silu4(pre);
                if (!rvalid) pre = make_float4(0.f, 0.f, 0.f, 0.f);
                split_tf32(pre.x, hi[4 * j4 + 0], lo[4 * j4 + 0]);
                split_tf32(pre.y, hi[4 * j4 + 1], lo[4 * j4 + 1]);
                split_tf32(pre.z, hi[4 * j4 + 2], lo[4 * j4 + 2]);
                split_tf32(pre.w, hi[4 * j4 + 3], lo[4 * j4 + 3]);
            }
            tmem_st16(lane_off + tA_hi + cc, hi);
            tmem_st16(lane_off + tA_lo + cc, lo);
        }
        wait_st();
        fence_before_sync();
        named_bar(bar_id, VT_GROUP);

        // ---- MMA 1 ----------------------------------------------------------------------------------
        if (u == 0) {
            fence_after_sync();
            issue_3xtf32<8, ((2 * B_LBO) >> 4)>(tD, tA_hi, tA_lo, dW2hi, dW2lo, idesc64, mbar);
        }
        __syncwarp();
        mbar_wait(mbar, 0);
        __syncwarp();
        fence_after_sync();

        // ---- stage 2: mv = SiLU(D + b2v) ------------------------------------------------------------
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int cc = cb + 16 * c;
            uint32_t d[16], hi[16], lo[16];
            tmem_ld16(lane_off + tD + cc, d);
            wait_ld();
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 bb = *reinterpret_cast<const float4*>(b2s + cc + 4 * j4);
                float4 m;
                m.x = silu(__uint_as_float(d[4 * j4 + 0]) + bb.x);
                m.y = silu(__uint_as_float(d[4 * j4 + 1]) + bb.y);
                m.z = silu(__uint_as_float(d[4 * j4 + 2]) + bb.z);
                m.w = silu(__uint_as_float(d[4 * j4 + 3]) + bb.w);
                if (need_feat) *reinterpret_cast<float4*>(tile_s + t * VT_ROW + cc + 4 * j4) = m;
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
        named_bar(bar_id, VT_GROUP);

        // ---- MMA 2 (both coordinate heads, N = 128) overlapped with the pools of mv ---------------------
        if (u == 0) {
            fence_after_sync();
            issue_3xtf32<8, ((2 * VT_LBO2) >> 4)>(tD, tA_hi, tA_lo, dWHhi, dWHlo, idesc128, mbar);
        }
        __syncwarp();
        if (need_feat) {
            const int c64 = u & 63, q4 = u >> 6;
            const float invC = 1.0f / (float)C;
            for (int n = q4; n < nvalid; n += 4) {          // mean over channels per node
                float s = 0.f;
                for (int c = 0; c < C; ++c) s += tile_s[(n * C + c) * VT_ROW + c64];
                a.agg_v[(size_t)(n0 + n) * H + c64] = s * invC;
            }
            if (single) {                                    // sum over nodes per channel
                for (int c = q4; c < C; c += 4) {
                    float s = 0.f;
                    for (int n = 0; n < nvalid; ++n) s += tile_s[(n * C + c) * VT_ROW + c64];
                    accH[c * H + c64] += s;
                }
            } else {
                for (int n = q4; n < nvalid; n += 4) {
                    float* dst = a.vsum + (size_t)sgraph[n] * K + 4 + 3 * C;
                    for (int c = 0; c < C; ++c) atomicAdd(dst + c * H + c64, tile_s[(n * C + c) * VT_ROW + c64]);
                }
            }
        }
        mbar_wait(mbar, 1);
        __syncwarp();
        fence_after_sync();

        // ---- stage 3: one head per column-half: φ = w3·SiLU(D_head + b_head) ---------------------------
        {
            float phi = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int cc = 64 * half + 16 * c;
                uint32_t d[16];
                tmem_ld16(lane_off + tD + cc, d);
                wait_ld();
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const float4 bb = *reinterpret_cast<const float4*>(bhs + cc + 4 * j4);
                    const float4 ww = *reinterpret_cast<const float4*>(w3s + cc + 4 * j4);
                    phi = fmaf(silu(__uint_as_float(d[4 * j4 + 0]) + bb.x), ww.x, phi);
                    phi = fmaf(silu(__uint_as_float(d[4 * j4 + 1]) + bb.y), ww.y, phi);
                    phi = fmaf(silu(__uint_as_float(d[4 * j4 + 2]) + bb.z), ww.z, phi);
                    phi = fmaf(silu(__uint_as_float(d[4 * j4 + 3]) + bb.w), ww.w, phi);
                }
            }
            phis[half * TILE_M + t] = phi;
        }
        fence_before_sync();
        named_bar(bar_id, VT_GROUP);
        // trans_v[node] = mean_c(−ΔX_c·φ_xv,c)
        for (int i = u; i < nvalid * 3; i += VT_GROUP) {
            const int n = i / 3, d = i - 3 * n;
            float s = 0.f;
            for (int c = 0; c < C; ++c) s = fmaf(-dXs[4 * (n * C + c) + d], phis[n * C + c], s);
            a.trans_v[(size_t)(n0 + n) * 4 + d] = s / (float)C;
        }
        // Σ_i ΔX_ic·φ_X,ic per graph, laid out [3][C]
        if (u < 3 * C) {
            const int d = u / C, c = u - d * C;
            const float* phx = phis + TILE_M;
            if (single) {
                float s = 0.f;
                for (int n = 0; n < nvalid; ++n) s = fmaf(dXs[4 * (n * C + c) + d], phx[n * C + c], s);
                accX[u] += s;
            } else {
                for (int n = 0; n < nvalid; ++n)
                    atomicAdd(a.vsum + (size_t)sgraph[n] * K + 4 + u, dXs[4 * (n * C + c) + d] * phx[n * C + c]);
            }
        }
        named_bar(bar_id, VT_GROUP);   // sgraph/dXs/phis/tile_s are rewritten by the next tile
    }
    flush(cur_graph);

    fence_before_sync();
    __syncthreads();
    if ((tid >> 5) == 0) tmem_dealloc(tbase, 512);
}

}  // namespace degnn

extern "C" int distegnn_virtual_layer_fwd_tf32(int64_t n_nodes, int n_graphs, int A, int C, int Na, unsigned flags,
                                          const int32_t* batch32, const float* x4, const float* Hn,
                                          const float* Xv, const float* G, const float* layer_params,
                                          float* agg_v, float* trans_v, float* vsum, void* stream) {
    using namespace degnn;
    if (int rc = check_dims(A, C, Na)) return rc;
    if (n_nodes == 0) return DISTEGNN_OK;
    DEGNN_CHECK_ARG(n_nodes > 0 && n_graphs > 0, "bad size");
    DEGNN_CHECK_ARG(batch32 && x4 && Hn && Xv && G && layer_params && trans_v && vsum, "null pointer");
    DEGNN_CHECK_ARG((flags & DISTEGNN_FLAG_LAST) || agg_v, "null agg_v");
    Layout L = make_layout(A, C, Na);
    VirtTcArgs a;
    a.N = n_nodes; a.B = n_graphs; a.C = C; a.flags = flags;
    a.batch = batch32; a.x4 = x4; a.Hn = Hn; a.Xv = Xv; a.G = G;
    a.w1r = layer_params + L.off[DISTEGNN_P_V_W1R];
    a.w2 = layer_params + L.off[DISTEGNN_P_V_W2];
    a.b2 = layer_params + L.off[DISTEGNN_P_V_B2];
    a.wxv = layer_params + L.off[DISTEGNN_P_V_WXV];
    a.bxv = layer_params + L.off[DISTEGNN_P_V_BXV];
    a.w3xv = layer_params + L.off[DISTEGNN_P_V_W3XV];
    a.wx = layer_params + L.off[DISTEGNN_P_V_WX];
    a.bx = layer_params + L.off[DISTEGNN_P_V_BX];
    a.w3x = layer_params + L.off[DISTEGNN_P_V_W3X];
    a.agg_v = agg_v; a.trans_v = trans_v; a.vsum = vsum;
    ensure_dynamic_smem((const void*)virtual_layer_tc_kernel, (int)VT_SMEM_BYTES);
    const int TN = TILE_M / C;
    const int64_t tiles = (n_nodes + TN - 1) / TN;
    int64_t grid = (tiles + 1) / 2;
    if (grid > sm_count()) grid = sm_count();
    virtual_layer_tc_kernel<<<(unsigned)grid, VT_THREADS, VT_SMEM_BYTES, (cudaStream_t)stream>>>(a);
    DEGNN_CHECK_LAUNCH();
    return DISTEGNN_OK;
}
