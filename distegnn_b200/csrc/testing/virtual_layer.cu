// Real<->virtual stage of one E_GCL_vel layer.  Replaces the virtual geometry, edge_mode_virtual and
// the virtual halves of coord_model_vel / coord_model_virtual / node_model / node_model_virtual
// (reference models/FastEGNN.py:252-253, 154-163, 180, 191-193, 207, 220-223) including the three
// global_mean_pool scatters (:193, :222) — here per-graph partial SUMS accumulated on chip and flushed
// once per CTA into the packed `vsum` buffer that is all-reduced across partitions.
//
// Rows of a tile are (node, channel) pairs: TN = 128 / C nodes per tile, row = n_local*C + c.  The
// reference materialises [N, 2H+1+C, C] and two [N, C, 64] tensors ("C times memory consumption",
// FastEGNN.py:266); here nothing of size N*C ever reaches HBM.
#include "common.cuh"

namespace degnn {

struct VirtArgs {
    int64_t N;
    int B, C;
    unsigned flags;
    const int32_t* batch;  // [N]
    const float* x4;       // [N,4]
    const float* Hn;       // [N,64]
    const float* Xv;       // [B,3,C]
    const float* G;        // [B,C,64]
    const float* w1r;      // [64]
    const float* w2; const float* b2;
    const float* wxv; const float* bxv; const float* w3xv;
    const float* wx; const float* bx; const float* w3x;
    float* agg_v;          // [N,64]
    float* trans_v;        // [N,4]
    float* vsum;           // [B,K]
};

constexpr int VMAXC = DISTEGNN_MAX_CHANNELS;
constexpr int VIRT_SMEM_FLOATS = TILE_M * LDA      // activation tile
                                 + 3 * H * H       // W2v, Wxv, WX
                                 + 8 * H           // b2, bxv, w3xv, bx, w3x, w1r (+2 spare)
                                 + TILE_M * 4      // ΔX per row
                                 + 2 * TILE_M      // φ_xv, φ_X per row
                                 + VMAXC * H       // Σ mv accumulators [C][64]
                                 + 4 * VMAXC       // Σ ΔX·φ_X accumulators [3][C]
                                 + TILE_M;         // graph id per local node (int)
constexpr size_t VIRT_SMEM_BYTES = VIRT_SMEM_FLOATS * sizeof(float);

__global__ void __launch_bounds__(NTHREADS, 2) virtual_layer_kernel(const VirtArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* As = smem;
    float* W2s = As + TILE_M * LDA;
    float* Wxvs = W2s + H * H;
    float* Wxs = Wxvs + H * H;
    float* b2s = Wxs + H * H;
    float* bxvs = b2s + H;
    float* w3xvs = bxvs + H;
    float* bxs = w3xvs + H;
    float* w3xs = bxs + H;
    float* w1rs = w3xs + H;
    float* dXs = w1rs + 3 * H;
    float* phixv = dXs + TILE_M * 4;
    float* phix = phixv + TILE_M;
    float* accH = phix + TILE_M;          // [C][64]
    float* accX = accH + VMAXC * H;       // [3][C]
    int* sgraph = reinterpret_cast<int*>(accX + 4 * VMAXC);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    const int C = a.C;
    const int K = 4 + 3 * C + H * C;
    const bool need_feat = !(a.flags & DISTEGNN_FLAG_LAST);
    const int TN = TILE_M / C;            // nodes per tile

    load_w64(W2s, a.w2, tid);
    load_w64(Wxvs, a.wxv, tid);
    load_w64(Wxs, a.wx, tid);
    if (tid < H) {
        b2s[tid] = a.b2[tid];
        bxvs[tid] = a.bxv[tid];
        w3xvs[tid] = a.w3xv[tid];
        bxs[tid] = a.bx[tid];
        w3xs[tid] = a.w3x[tid];
        w1rs[tid] = a.w1r[tid];
    }
    for (int i = tid; i < VMAXC * H + 4 * VMAXC; i += NTHREADS) accH[i] = 0.f;   // accH and accX
    int cur_graph = -1;   // graph whose sums sit in accH/accX (uniform across the CTA)
    __syncthreads();

    auto flush = [&](int g) {   // all threads; adds the CTA-local sums of graph g into vsum
        if (g >= 0) {
            float* dst = a.vsum + (size_t)g * K;
            if (need_feat)
                for (int i = tid; i < C * H; i += NTHREADS) {
                    atomicAdd(dst + 4 + 3 * C + i, accH[i]);
                    accH[i] = 0.f;
                }
            if (tid < 3 * C) {
                atomicAdd(dst + 4 + tid, accX[tid]);
                accX[tid] = 0.f;
            }
        }
    };

    const int64_t num_tiles = (a.N + TN - 1) / TN;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t n0 = tile * TN;
        const int nvalid = (int)min((int64_t)TN, a.N - n0);
        const int rows = nvalid * C;
        if (tid < TN) sgraph[tid] = (tid < nvalid) ? __ldg(a.batch + n0 + tid) : -1;
        __syncthreads();
        const int g_first = sgraph[0], g_last = sgraph[nvalid - 1];
        const bool single = (g_first == g_last);
        if (single && g_first != cur_graph) {
            flush(cur_graph);
            cur_graph = g_first;
            __syncthreads();
        }

        // ---- first layer: half-warp per (node,channel) row ----
        {
            const int l = lane & 15;
            const float4 wr4 = *reinterpret_cast<const float4*>(w1rs + 4 * l);
#pragma unroll 4
            for (int it = 0; it < 8; ++it) {
                const int r = 16 * warp + 2 * it + (lane >> 4);
                float4 pre = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < rows) {
                    const int nl = r / C, c = r - nl * C;
                    const int g = sgraph[nl];
                    const int64_t node = n0 + nl;
                    float4 xi = ldg4(a.x4 + (size_t)node * 4);
                    const float* Xg = a.Xv + (size_t)g * 3 * C;
                    float dx = __ldg(Xg + c) - xi.x;
                    float dy = __ldg(Xg + C + c) - xi.y;
                    float dz = __ldg(Xg + 2 * C + c) - xi.z;
                    float vr = sqrtf(dx * dx + dy * dy + dz * dz);
                    float4 hn = ldg4(a.Hn + (size_t)node * H + 4 * l);
                    float4 gg = ldg4(a.G + ((size_t)g * C + c) * H + 4 * l);
                    pre = silu4(fma4(vr, wr4, add4(hn, gg)));
                    if (l == 0) *reinterpret_cast<float4*>(dXs + 4 * r) = make_float4(dx, dy, dz, 0.f);
                }
                *reinterpret_cast<float4*>(As + r * LDA + 4 * l) = pre;
            }
        }
        __syncthreads();

        // ---- mv = SiLU(W2v·a1 + b2v) ----
        float acc[8][4];
        zero_acc(acc);
        gemm_tile(acc, As, W2s, ty, tx);
        __syncthreads();
        bias_silu_to_tile(acc, *reinterpret_cast<const float4*>(b2s + 4 * tx), As, ty, tx);
        __syncthreads();

        // ---- per-node mean over channels, per-graph sum over nodes ----
        if (need_feat) {
            const int c64 = tid & 63, q = tid >> 6;
            const float invC = 1.0f / (float)C;
            for (int nl = q; nl < nvalid; nl += 4) {
                float s = 0.f;
                for (int c = 0; c < C; ++c) s += As[(nl * C + c) * LDA + c64];
                a.agg_v[(size_t)(n0 + nl) * H + c64] = s * invC;
            }
            if (single) {
                for (int c = q; c < C; c += 4) {
                    float s = 0.f;
                    for (int nl = 0; nl < nvalid; ++nl) s += As[(nl * C + c) * LDA + c64];
                    accH[c * H + c64] += s;
                }
            } else {
                for (int nl = q; nl < nvalid; nl += 4) {
                    float* dst = a.vsum + (size_t)sgraph[nl] * K + 4 + 3 * C;
                    for (int c = 0; c < C; ++c) atomicAdd(dst + c * H + c64, As[(nl * C + c) * LDA + c64]);
                }
            }
        }

        // ---- φ_xv and φ_X heads ----
        zero_acc(acc);
        gemm_tile(acc, As, Wxvs, ty, tx);
        head_dot_to_smem(acc, *reinterpret_cast<const float4*>(bxvs + 4 * tx),
                         *reinterpret_cast<const float4*>(w3xvs + 4 * tx), phixv, ty, tx);
        zero_acc(acc);
        gemm_tile(acc, As, Wxs, ty, tx);
        head_dot_to_smem(acc, *reinterpret_cast<const float4*>(bxs + 4 * tx),
                         *reinterpret_cast<const float4*>(w3xs + 4 * tx), phix, ty, tx);
        __syncthreads();

        // trans_v[node] = mean_c(−ΔX_c·φ_xv,c)
        for (int i = tid; i < nvalid * 3; i += NTHREADS) {
            const int nl = i / 3, d = i - 3 * nl;
            float s = 0.f;
            for (int c = 0; c < C; ++c) s = fmaf(-dXs[4 * (nl * C + c) + d], phixv[nl * C + c], s);
            a.trans_v[(size_t)(n0 + nl) * 4 + d] = s / (float)C;
        }
        // Σ_i ΔX_ic·φ_X,ic per graph, laid out [3][C]
        if (tid < 3 * C) {
            const int d = tid / C, c = tid - d * C;
            if (single) {
                float s = 0.f;
                for (int nl = 0; nl < nvalid; ++nl) s = fmaf(dXs[4 * (nl * C + c) + d], phix[nl * C + c], s);
                accX[tid] += s;
            } else {
                for (int nl = 0; nl < nvalid; ++nl)
                    atomicAdd(a.vsum + (size_t)sgraph[nl] * K + 4 + tid,
                              dXs[4 * (nl * C + c) + d] * phix[nl * C + c]);
            }
        }
        __syncthreads();
    }
    flush(cur_graph);
}

}  // namespace degnn

extern "C" int distegnn_virtual_layer_fwd_simt(int64_t n_nodes, int n_graphs, int A, int C, int Na,
                                          unsigned flags, const int32_t* batch32, const float* x4,
                                          const float* Hn, const float* Xv, const float* G,
                                          const float* layer_params, float* agg_v, float* trans_v,
                                          float* vsum, void* stream) {
    using namespace degnn;
    if (int rc = check_dims(A, C, Na)) return rc;
    if (n_nodes == 0) return DISTEGNN_OK;
    DEGNN_CHECK_ARG(n_nodes > 0 && n_graphs > 0, "bad size");
    DEGNN_CHECK_ARG(batch32 && x4 && Hn && Xv && G && layer_params && trans_v && vsum, "null pointer");
    DEGNN_CHECK_ARG((flags & DISTEGNN_FLAG_LAST) || agg_v, "null agg_v");
    Layout L = make_layout(A, C, Na);
    VirtArgs a;
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

    ensure_dynamic_smem((const void*)virtual_layer_kernel, (int)VIRT_SMEM_BYTES);
    const int TN = TILE_M / C;
    int64_t tiles = (n_nodes + TN - 1) / TN;
    int64_t grid = (int64_t)sm_count() * 2;
    if (grid > tiles) grid = tiles;
    virtual_layer_kernel<<<(unsigned)grid, NTHREADS, VIRT_SMEM_BYTES, (cudaStream_t)stream>>>(a);
    DEGNN_CHECK_LAUNCH();
    return DISTEGNN_OK;
}
