// Backward of the real<->virtual stage — fp32 FMA on the CUDA cores: the first, correctness-first kernel,
// now behind distegnn_virtual_layer_bwd_simt as the twin of the tensor-core kernel (virtual_layer_bwd_tc.cu).
// Differentiates what distegnn_virtual_layer_fwd computes (reference models/FastEGNN.py:252-253 virtual geometry,
// :154-163 edge_mode_virtual, :180 / :191-193 / :207 / :220-223 virtual halves of the coordinate and feature models and
// the global_mean_pool scatters; in the reference: autograd through [N,2H+1+C,C] and several [N,C,64] tensors).
// Per row (node i of graph b, channel c), recomputed tile by tile (rows of a tile = 128/C consecutive nodes x C):
//     ΔX = Xv[b,:,c] − x_i,  vr = ‖ΔX‖,  z1 = Hn_i + G[b,c] + w_vr·vr,  a1 = SiLU(z1),  z2 = W2v·a1 + b2v,
//     mv = SiLU(z2),  φ_xv = w3xv·SiLU(Wxv·mv + bxv),  φ_X = w3x·SiLU(Wx·mv + bx)
//     agg_v[i] = mean_c mv,  trans_v[i] = mean_c(−ΔX·φ_xv),  vsum[b, 4+d·C+c] += ΔX_d·φ_X,  vsum[b, 4+3C+c·64+n] += mv_n
// Upstream gradients: g_agg_v [N,64], g_trans_v [N,4], g_vsum [B,K] (already all-reduced over the partitions).
// Outputs: g_Hn [N,64] and g_xv [N,4] are WRITTEN (one tile per node); g_G [B,C,64], g_Xv [B,3,C] and the parameter
// gradients (fields V_W1R, V_W2, V_B2, V_WXV, V_BXV, V_W3XV, V_WX, V_BX, V_W3X of a parameter-layout buffer) are
// ACCUMULATED (+=).  The transposed 64x64 weight matrices the data-gradient GEMMs need are passed in `wT`
// ([3][64][64]: W2vᵀ, Wxvᵀ, Wxᵀ in the same k-major convention), built by the caller.
#include "bwd_common.cuh"
#include "common.cuh"

namespace degnn {

struct VirtBwdArgs {
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
    const float* wT;          // [3][64][64]
    const float* g_aggv;      // [N,64] (null with FLAG_LAST)
    const float* g_transv;    // [N,4]
    const float* g_vsum;      // [B,K]
    float* g_Hn;              // [N,64] =
    float* g_xv;              // [N,4]  =
    float* g_G;               // [B,C,64] +=
    float* g_Xv;              // [B,3,C]  +=
    float* g_w1r; float* g_w2; float* g_b2; float* g_wxv; float* g_bxv; float* g_w3xv;
    float* g_wx; float* g_bx; float* g_w3x;
};

constexpr int VB_MAXC = DISTEGNN_MAX_CHANNELS;
constexpr int VB_SMEM_FLOATS = 4 * TILE_M * LDA        // Z1, Z2, Wt, Gt
                               + 6 * H + 6 * H         // w1r, b2, bxv, w3xv, bx, w3x and their gradient accumulators
                               + VB_MAXC * H           // Σ_i g_z1 per channel (-> g_G)
                               + 4 * VB_MAXC           // Σ_i gΔX per channel (-> g_Xv)
                               + TILE_M * 4            // ΔX xyz, vr
                               + TILE_M * 4            // gΔX xyz
                               + 5 * TILE_M            // gφ_xv, gφ_X, φ_xv, φ_X, g_vr
                               + TILE_M;               // graph id per local node (int)
constexpr size_t VB_SMEM_BYTES = VB_SMEM_FLOATS * sizeof(float);

__global__ void __launch_bounds__(NTHREADS, 1) virtual_layer_bwd_kernel(const VirtBwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* Z1 = smem;
    float* Z2 = Z1 + TILE_M * LDA;
    float* Wt = Z2 + TILE_M * LDA;
    float* Gt = Wt + TILE_M * LDA;
    float* w1rs = Gt + TILE_M * LDA;
    float* b2s = w1rs + H;
    float* bxvs = b2s + H;
    float* w3xvs = bxvs + H;
    float* bxs = w3xvs + H;
    float* w3xs = bxs + H;
    float* gw1r = w3xs + H;
    float* gb2 = gw1r + H;
    float* gbxv = gb2 + H;
    float* gw3xv = gbxv + H;
    float* gbx = gw3xv + H;
    float* gw3x = gbx + H;
    float* accG = gw3x + H;                    // [C][64]
    float* accX = accG + VB_MAXC * H;          // [3][C] (row pitch VB_MAXC)
    float* dXs = accX + 4 * VB_MAXC;           // [128][4]
    float* gdX = dXs + TILE_M * 4;             // [128][4]
    float* gpxv = gdX + TILE_M * 4;
    float* gpx = gpxv + TILE_M;
    float* pxv = gpx + TILE_M;
    float* px = pxv + TILE_M;
    float* gvr = px + TILE_M;
    int* sgraph = reinterpret_cast<int*>(gvr + TILE_M);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    const int C = a.C;
    const int K = 4 + 3 * C + H * C;
    const int TN = TILE_M / C;
    const float invC = 1.0f / (float)C;
    const bool need_feat = !(a.flags & DISTEGNN_FLAG_LAST) && a.g_aggv != nullptr;
    const float* W2T = a.wT;
    const float* WxvT = a.wT + H * H;
    const float* WxT = a.wT + 2 * H * H;

    if (tid < H) {
        w1rs[tid] = a.w1r[tid];
        b2s[tid] = a.b2[tid];
        bxvs[tid] = a.bxv[tid];
        w3xvs[tid] = a.w3xv[tid];
        bxs[tid] = a.bx[tid];
        w3xs[tid] = a.w3x[tid];
    }
    for (int i = tid; i < 6 * H + VB_MAXC * H + 4 * VB_MAXC; i += NTHREADS) gw1r[i] = 0.f;
    float gW2[4][4], gWxv[4][4], gWx[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) gW2[i][j] = gWxv[i][j] = gWx[i][j] = 0.f;
    __syncthreads();
    const float4 b2v = *reinterpret_cast<const float4*>(b2s + 4 * tx);
    const float4 wrv = *reinterpret_cast<const float4*>(w1rs + 4 * tx);

    int cur_graph = -1;
    auto flush = [&](int g) {                  // all threads; caller synchronises
        if (g >= 0) {
            for (int i = tid; i < C * H; i += NTHREADS) {
                atomicAdd(a.g_G + (size_t)g * C * H + i, accG[i]);
                accG[i] = 0.f;
            }
            if (tid < 3 * C) {
                const int d = tid / C, c = tid - d * C;
                atomicAdd(a.g_Xv + (size_t)g * 3 * C + tid, accX[d * VB_MAXC + c]);
                accX[d * VB_MAXC + c] = 0.f;
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
        const int g_first = sgraph[0];
        const bool single = (g_first == sgraph[nvalid - 1]);
        if (single && g_first != cur_graph) {
            flush(cur_graph);
            cur_graph = g_first;
            __syncthreads();
        }

        // ---- 1. first layer (half-warp per row): Z1 = z1, Wt = a1, geometry, upstream scalars -----------------------
        {
            const int l = lane & 15;
            const float4 wr4 = *reinterpret_cast<const float4*>(w1rs + 4 * l);
#pragma unroll 2
            for (int it = 0; it < 8; ++it) {
                const int rr = 16 * warp + 2 * it + (lane >> 4);
                float4 pre = make_float4(0.f, 0.f, 0.f, 0.f);
                if (rr < rows) {
                    const int nl = rr / C, c = rr - nl * C;
                    const int g = sgraph[nl];
                    const size_t node = (size_t)(n0 + nl);
                    const float4 xi = ldg4(a.x4 + node * 4);
                    const float* Xg = a.Xv + (size_t)g * 3 * C;
                    const float dx = __ldg(Xg + c) - xi.x, dy = __ldg(Xg + C + c) - xi.y, dz = __ldg(Xg + 2 * C + c) - xi.z;
                    const float vr = sqrtf(dx * dx + dy * dy + dz * dz);
                    pre = fma4(vr, wr4, add4(ldg4(a.Hn + node * H + 4 * l), ldg4(a.G + ((size_t)g * C + c) * H + 4 * l)));
                    if (l == 0) {
                        const float4 gt = ldg4(a.g_transv + node * 4);
                        const float* gv = a.g_vsum + (size_t)g * K + 4;
                        *reinterpret_cast<float4*>(dXs + 4 * rr) = make_float4(dx, dy, dz, vr);
                        gpxv[rr] = -(gt.x * dx + gt.y * dy + gt.z * dz) * invC;
                        gpx[rr] = __ldg(gv + c) * dx + __ldg(gv + C + c) * dy + __ldg(gv + 2 * C + c) * dz;
                    }
                } else if (l == 0) {
                    *reinterpret_cast<float4*>(dXs + 4 * rr) = make_float4(0.f, 0.f, 0.f, 0.f);
                    gpxv[rr] = 0.f;
                    gpx[rr] = 0.f;
                }
                *reinterpret_cast<float4*>(Z1 + rr * LDA + 4 * l) = pre;
                *reinterpret_cast<float4*>(Wt + rr * LDA + 4 * l) = silu4(pre);
            }
        }
        __syncthreads();

        // ---- 2. z2 = a1·W2v + b2v -> Z2;  Wt = mv ---------------------------------------------------------------------
        float acc[8][4], accm[8][4];
        zero_acc(acc);
        gemm_tile_g(acc, Wt, a.w2, ty, tx);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 z = make_float4(acc[i][0] + b2v.x, acc[i][1] + b2v.y, acc[i][2] + b2v.z, acc[i][3] + b2v.w);
            *reinterpret_cast<float4*>(Z2 + (ty + 16 * i) * LDA + 4 * tx) = z;
            *reinterpret_cast<float4*>(Wt + (ty + 16 * i) * LDA + 4 * tx) = silu4(z);
        }
        __syncthreads();

        // ---- 3. the two coordinate heads: φ, g_z(head) -> Gt, weight gradients, g_mv accumulated in accm --------------
        zero_acc(accm);
        auto head = [&](const float* wg, const float* wTg, const float* bs, const float* w3s_, const float* gps, float* phs,
                        float* gw3acc, float* gbacc, float (&gWh)[4][4]) {
            zero_acc(acc);
            gemm_tile_g(acc, Wt, wg, ty, tx);
            const float4 bv = *reinterpret_cast<const float4*>(bs + 4 * tx);
            const float4 w3v = *reinterpret_cast<const float4*>(w3s_ + 4 * tx);
            float gw[4] = {0.f, 0.f, 0.f, 0.f}, gb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int e = ty + 16 * i;
                const float gp = gps[e];
                const float zc[4] = {acc[i][0] + bv.x, acc[i][1] + bv.y, acc[i][2] + bv.z, acc[i][3] + bv.w};
                const float w3a[4] = {w3v.x, w3v.y, w3v.z, w3v.w};
                float ph = 0.f, g[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float s = sigmoid_f(zc[j]);
                    const float ac = zc[j] * s;
                    ph = fmaf(ac, w3a[j], ph);
                    gw[j] = fmaf(gp, ac, gw[j]);
                    g[j] = gp * w3a[j] * (s * fmaf(zc[j], 1.0f - s, 1.0f));
                    gb[j] += g[j];
                }
                ph += __shfl_xor_sync(FULL, ph, 1);
                ph += __shfl_xor_sync(FULL, ph, 2);
                ph += __shfl_xor_sync(FULL, ph, 4);
                ph += __shfl_xor_sync(FULL, ph, 8);
                if (tx == 0) phs[e] = ph;
                *reinterpret_cast<float4*>(Gt + e * LDA + 4 * tx) = make_float4(g[0], g[1], g[2], g[3]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                atomicAdd(gw3acc + 4 * tx + j, gw[j]);
                atomicAdd(gbacc + 4 * tx + j, gb[j]);
            }
            __syncthreads();
            wgrad_tile(gWh, Gt, Wt, tid);
            gemm_tile_g(accm, Gt, wTg, ty, tx);
            __syncthreads();                   // Gt fully read before the next head overwrites it
        };
        head(a.wxv, WxvT, bxvs, w3xvs, gpxv, pxv, gw3xv, gbxv, gWxv);
        head(a.wx, WxT, bxs, w3xs, gpx, px, gw3x, gbx, gWx);

        // ---- 4. g_mv (+ upstream) -> g_z2 = g_mv ⊙ SiLU'(z2) -> Gt;  Wt = a1 again --------------------------------------
        {
            float gb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int e = ty + 16 * i;
                float4 gm = make_float4(accm[i][0], accm[i][1], accm[i][2], accm[i][3]);
                if (need_feat && e < rows) {
                    const int nl = e / C, c = e - nl * C;
                    const float4 ga = ldg4(a.g_aggv + (size_t)(n0 + nl) * H + 4 * tx);
                    const float* gsp = a.g_vsum + (size_t)sgraph[nl] * K + 4 + 3 * C + c * H + 4 * tx;   // K is odd in
                    const float4 gs = make_float4(__ldg(gsp), __ldg(gsp + 1), __ldg(gsp + 2), __ldg(gsp + 3));   // general
                    gm.x += fmaf(ga.x, invC, gs.x); gm.y += fmaf(ga.y, invC, gs.y);
                    gm.z += fmaf(ga.z, invC, gs.z); gm.w += fmaf(ga.w, invC, gs.w);
                }
                const float4 z = *reinterpret_cast<const float4*>(Z2 + e * LDA + 4 * tx);
                const float4 g = make_float4(gm.x * dsilu(z.x), gm.y * dsilu(z.y), gm.z * dsilu(z.z), gm.w * dsilu(z.w));
                gb[0] += g.x; gb[1] += g.y; gb[2] += g.z; gb[3] += g.w;
                *reinterpret_cast<float4*>(Gt + e * LDA + 4 * tx) = g;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(gb2 + 4 * tx + j, gb[j]);
        }
        silu_tile(Wt, Z1, tid);
        __syncthreads();

        // ---- 5. g_W2v += g_z2ᵀ·a1;  g_z1 = (g_z2·W2v) ⊙ SiLU'(z1) -> Z2 tile;  g_vr, g_w_vr -----------------------------
        wgrad_tile(gW2, Gt, Wt, tid);
        zero_acc(acc);
        gemm_tile_g(acc, Gt, W2T, ty, tx);
        {
            float gwr[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int e = ty + 16 * i;
                const float4 z = *reinterpret_cast<const float4*>(Z1 + e * LDA + 4 * tx);
                const float4 g = make_float4(acc[i][0] * dsilu(z.x), acc[i][1] * dsilu(z.y), acc[i][2] * dsilu(z.z),
                                             acc[i][3] * dsilu(z.w));
                float gr = g.x * wrv.x + g.y * wrv.y + g.z * wrv.z + g.w * wrv.w;
                gr += __shfl_xor_sync(FULL, gr, 1);
                gr += __shfl_xor_sync(FULL, gr, 2);
                gr += __shfl_xor_sync(FULL, gr, 4);
                gr += __shfl_xor_sync(FULL, gr, 8);
                if (tx == 0) gvr[e] = gr;
                const float vr = dXs[4 * e + 3];
                gwr[0] = fmaf(g.x, vr, gwr[0]); gwr[1] = fmaf(g.y, vr, gwr[1]);
                gwr[2] = fmaf(g.z, vr, gwr[2]); gwr[3] = fmaf(g.w, vr, gwr[3]);
                *reinterpret_cast<float4*>(Z2 + e * LDA + 4 * tx) = g;          // Z2 is dead: reuse it for g_z1
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(gw1r + 4 * tx + j, gwr[j]);
        }
        __syncthreads();

        // ---- 6. reductions of g_z1 and the geometry gradient ------------------------------------------------------------
        {
            const int c64 = tid & 63, q = tid >> 6;
            for (int n = q; n < nvalid; n += 4) {           // g_Hn[node] = Σ_c g_z1
                float s = 0.f;
                for (int c = 0; c < C; ++c) s += Z2[(n * C + c) * LDA + c64];
                a.g_Hn[(size_t)(n0 + n) * H + c64] = s;
            }
            if (single) {                                    // Σ_i g_z1 per channel -> g_G
                for (int c = q; c < C; c += 4) {
                    float s = 0.f;
                    for (int n = 0; n < nvalid; ++n) s += Z2[(n * C + c) * LDA + c64];
                    accG[c * H + c64] += s;
                }
            } else {
                for (int n = q; n < nvalid; n += 4)
                    for (int c = 0; c < C; ++c)
                        atomicAdd(a.g_G + ((size_t)sgraph[n] * C + c) * H + c64, Z2[(n * C + c) * LDA + c64]);
            }
        }
        if (tid < rows) {                                    // gΔX = −g_trans_v·φ_xv/C + g_vsum·φ_X + g_vr·ΔX/vr
            const int nl = tid / C, c = tid - nl * C;
            const float4 d = *reinterpret_cast<const float4*>(dXs + 4 * tid);
            const float4 gt = ldg4(a.g_transv + (size_t)(n0 + nl) * 4);
            const float* gv = a.g_vsum + (size_t)sgraph[nl] * K + 4;
            const float s1 = -pxv[tid] * invC, s2 = px[tid], s3 = d.w > 0.f ? gvr[tid] / d.w : 0.f;
            *reinterpret_cast<float4*>(gdX + 4 * tid) =
                make_float4(fmaf(gt.x, s1, fmaf(__ldg(gv + c), s2, s3 * d.x)), fmaf(gt.y, s1, fmaf(__ldg(gv + C + c), s2, s3 * d.y)),
                            fmaf(gt.z, s1, fmaf(__ldg(gv + 2 * C + c), s2, s3 * d.z)), 0.f);
        }
        __syncthreads();
        if (tid < nvalid) {                                  // g_x (virtual part) = −Σ_c gΔX
            float sx = 0.f, sy = 0.f, sz = 0.f;
            for (int c = 0; c < C; ++c) {
                const float4 g = *reinterpret_cast<const float4*>(gdX + 4 * (tid * C + c));
                sx += g.x; sy += g.y; sz += g.z;
            }
            *reinterpret_cast<float4*>(a.g_xv + (size_t)(n0 + tid) * 4) = make_float4(-sx, -sy, -sz, 0.f);
        }
        if (tid >= 128 && tid < 128 + 3 * C) {               // g_Xv[b,d,c] += Σ_i gΔX_d
            const int k = tid - 128, d = k / C, c = k - d * C;
            if (single) {
                float s = 0.f;
                for (int n = 0; n < nvalid; ++n) s += gdX[4 * (n * C + c) + d];
                accX[d * VB_MAXC + c] += s;
            } else {
                for (int n = 0; n < nvalid; ++n) atomicAdd(a.g_Xv + (size_t)sgraph[n] * 3 * C + k, gdX[4 * (n * C + c) + d]);
            }
        }
        __syncthreads();                                     // tile buffers are rewritten by the next iteration
    }
    flush(cur_graph);

    wgrad_flush(a.g_w2, gW2, tid);
    wgrad_flush(a.g_wxv, gWxv, tid);
    wgrad_flush(a.g_wx, gWx, tid);
    if (tid < H) {
        atomicAdd(a.g_w1r + tid, gw1r[tid]);
        atomicAdd(a.g_b2 + tid, gb2[tid]);
        atomicAdd(a.g_bxv + tid, gbxv[tid]);
        atomicAdd(a.g_w3xv + tid, gw3xv[tid]);
        atomicAdd(a.g_bx + tid, gbx[tid]);
        atomicAdd(a.g_w3x + tid, gw3x[tid]);
    }
}

}  // namespace degnn

extern "C" int distegnn_virtual_layer_bwd_simt(int64_t n_nodes, int n_graphs, int A, int C, int Na, unsigned flags,
                                          const int32_t* batch32, const float* x4, const float* Hn, const float* Xv,
                                          const float* G, const float* layer_params, const float* wT,
                                          const float* g_agg_v, const float* g_trans_v, const float* g_vsum,
                                          float* g_Hn, float* g_xv, float* g_G, float* g_Xv, float* g_layer_params,
                                          void* stream) {
    using namespace degnn;
    if (int rc = check_dims(A, C, Na)) return rc;
    if (n_nodes == 0) return DISTEGNN_OK;
    DEGNN_CHECK_ARG(n_nodes > 0 && n_graphs > 0, "bad size");
    DEGNN_CHECK_ARG(batch32 && x4 && Hn && Xv && G && layer_params && wT && g_trans_v && g_vsum && g_Hn && g_xv && g_G &&
                        g_Xv && g_layer_params,
                    "null pointer");
    Layout L = make_layout(A, C, Na);
    VirtBwdArgs a;
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
    a.wT = wT;
    a.g_aggv = g_agg_v; a.g_transv = g_trans_v; a.g_vsum = g_vsum;
    a.g_Hn = g_Hn; a.g_xv = g_xv; a.g_G = g_G; a.g_Xv = g_Xv;
    a.g_w1r = g_layer_params + L.off[DISTEGNN_P_V_W1R];
    a.g_w2 = g_layer_params + L.off[DISTEGNN_P_V_W2];
    a.g_b2 = g_layer_params + L.off[DISTEGNN_P_V_B2];
    a.g_wxv = g_layer_params + L.off[DISTEGNN_P_V_WXV];
    a.g_bxv = g_layer_params + L.off[DISTEGNN_P_V_BXV];
    a.g_w3xv = g_layer_params + L.off[DISTEGNN_P_V_W3XV];
    a.g_wx = g_layer_params + L.off[DISTEGNN_P_V_WX];
    a.g_bx = g_layer_params + L.off[DISTEGNN_P_V_BX];
    a.g_w3x = g_layer_params + L.off[DISTEGNN_P_V_W3X];
    ensure_dynamic_smem((const void*)virtual_layer_bwd_kernel, (int)VB_SMEM_BYTES);
    const int TN = TILE_M / C;
    const int64_t tiles = (n_nodes + TN - 1) / TN;
    int64_t grid = sm_count();
    if (grid > tiles) grid = tiles;
    virtual_layer_bwd_kernel<<<(unsigned)grid, NTHREADS, VB_SMEM_BYTES, (cudaStream_t)stream>>>(a);
    DEGNN_CHECK_LAUNCH();
    return DISTEGNN_OK;
}
