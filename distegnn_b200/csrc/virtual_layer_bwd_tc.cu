// Backward of the real<->virtual stage on the 5th-gen tensor cores — production kernel behind
// distegnn_virtual_layer_bwd (the fp32-FMA kernel of virtual_layer_bwd.cu is kept as distegnn_virtual_layer_bwd_simt).
// Same contract and math as virtual_layer_bwd.cu (reference: autograd through models/FastEGNN.py:154-163, 180, 191-193,
// 207, 220-223, 252-253).  Per 128-row tile (rows = (node, channel)) the SIX row-wise tile GEMMs — recompute z2 = a1·W2vᵀ,
// zxv = mv·Wxvᵀ, zx = mv·Wxᵀ; data gradients g_mv = g_zxv·Wxv + g_zx·Wx (two MMAs into ONE accumulator, both rows encoded
// with a common scale), g_a1 = g_z2·W2v — run as tcgen05.mma kind::f16 with the fp16 2-term split (tc16.cuh), A written to
// TMEM by the thread that owns the row; the three weight-gradient GEMMs stay on the CUDA cores (see edge_layer_bwd_tc.cu).
//
// Shared memory: six 64x64 B operands (hi + lo = 16 KB each) do not fit next to the two fp32 row tiles of two tile groups,
// so the weights are NOT resident: distegnn_virtual_bwd_prepare writes them once per call as fp16 hi/lo IMAGES already in
// the shared-memory operand layout, and every tile group streams them through two 16 KB slots with one TMA bulk copy per
// matrix, two matrices ahead of their use (mbarrier per slot; the order W2v, Wxv, Wx, Wxvᵀ, Wxᵀ, W2vᵀ repeats every tile).
// One CTA per SM, 256 threads = 2 tile groups of 4 warps; thread r owns row r and holds whole 64-wide rows in registers.
// TMEM per group (256 columns): A_hi 32 | A_lo 32 | D 64 | z1 64 | z2 64.  Every row is encoded with its own power-of-two
// scale (gradient rows span many orders of magnitude).
#include <cuda_fp16.h>

#include "bwd_common.cuh"
#include "bwd_tc_common.cuh"
#include "common.cuh"
#include "tc16.cuh"
#include "umma.cuh"

namespace degnn {

struct VirtBwdTcArgs {
    int64_t N;
    int B, C;
    unsigned flags;
    const int32_t* batch;
    const float* x4;
    const float* Hn;
    const float* Xv;
    const float* G;
    const float* w1r;
    const float* b2; const float* bxv; const float* w3xv; const float* bx; const float* w3x;
    const __half* wimg;       // [6][hi 4096 | lo 4096] operand images: W2v, Wxv, Wx, Wxvᵀ, Wxᵀ, W2vᵀ
    const float* g_aggv;
    const float* g_transv;
    const float* g_vsum;
    float* g_Hn;
    float* g_xv;
    float* g_G;
    float* g_Xv;
    float* g_w1r; float* g_w2; float* g_b2; float* g_wxv; float* g_bxv; float* g_w3xv;
    float* g_wx; float* g_bx; float* g_w3x;
};

constexpr int VT_THREADS = 256, VT_GROUPS = 2, VT_GROUP = 128;
constexpr int VT_MAXC = DISTEGNN_MAX_CHANNELS;
constexpr int VT_IMG = 2 * 64 * 64;                                 // halfs per matrix image (hi + lo)
constexpr int VT_SMEM_BYTES = VT_GROUPS * 2 * VT_IMG * 2            // two weight slots per group
                              + VT_GROUPS * 2 * TILE_M * LDA * 4    // gradient tile + activation tile per group
                              + 6 * H * 4 + 6 * H * 4               // w1r, b2, bxv, w3xv, bx, w3x + gradient accumulators
                              + VT_GROUPS * VT_MAXC * H * 4         // Σ_i g_z1 per channel (-> g_G)
                              + VT_GROUPS * 4 * VT_MAXC * 4         // Σ_i gΔX per channel (-> g_Xv)
                              + VT_GROUPS * TILE_M * 4 * 4          // gΔX per row
                              + VT_GROUPS * TILE_M * 4              // ‖ΔX‖ per row
                              + VT_GROUPS * TILE_M * 4              // graph id per local node
                              + 256;                                // mbarriers + tmem base
constexpr uint32_t VT_LBO = 1024;

// fp16 hi/lo images of the six B operands in the shared-memory layout (K-major, no swizzle): element (n,k) of an image at
// (k/8)*512 + (n/8)*64 + (n%8)*8 + k%8 halfs; forward matrices B[n][k] = W[n][k] = w_kmajor[k*64+n], transposed ones
// B[n][k] = W[k][n] = w_kmajor[n*64+k].
__global__ void virtual_bwd_images_kernel(const float* w2, const float* wxv, const float* wx, __half* img) {
    const int m = blockIdx.x;                    // 0..5
    const float* src = (m == 0 || m == 5) ? w2 : ((m == 1 || m == 3) ? wxv : wx);
    const bool transposed = m >= 3;
    __half* hi = img + (size_t)m * VT_IMG;
    __half* lo = hi + 64 * 64;
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
        const int n = i >> 6, k = i & 63;
        const float w = transposed ? src[n * 64 + k] : src[k * 64 + n];
        const __half h = __float2half_rn(w);
        const uint32_t o = (uint32_t)(k >> 3) * 512u + (uint32_t)(n >> 3) * 64u + (uint32_t)(n & 7) * 8u + (k & 7);
        hi[o] = h;
        lo[o] = __float2half_rn(w - __half2float(h));
    }
}

__global__ void __launch_bounds__(VT_THREADS, 1) virtual_layer_bwd_tc_kernel(const VirtBwdTcArgs a) {
    using namespace umma;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __half* slots = reinterpret_cast<__half*>(smem_raw);                     // [2 groups][2 slots][VT_IMG]
    float* tiles = reinterpret_cast<float*>(slots + VT_GROUPS * 2 * VT_IMG); // [2 groups][G tile | Act tile]
    float* w1rs = tiles + VT_GROUPS * 2 * TILE_M * LDA;
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
    float* accG_all = gw3x + H;                                              // [2][C][64]
    float* accX_all = accG_all + VT_GROUPS * VT_MAXC * H;                    // [2][3][VT_MAXC] (pitch VT_MAXC, 4 rows)
    float* gdX_all = accX_all + VT_GROUPS * 4 * VT_MAXC;                     // [2][128][4]
    float* vrs_all = gdX_all + VT_GROUPS * TILE_M * 4;                       // [2][128]
    int* sgraph_all = reinterpret_cast<int*>(vrs_all + VT_GROUPS * TILE_M);  // [2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sgraph_all + VT_GROUPS * TILE_M);   // [2][mma, slot0, slot1]
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 3 * VT_GROUPS);

    const int tid = threadIdx.x;
    const int grp = tid >> 7, t = tid & 127, lane = tid & 31, wq = (tid >> 5) & 3;
    const int C = a.C;
    const int K = 4 + 3 * C + H * C;
    const int TN = TILE_M / C;
    const float invC = 1.0f / (float)C;
    const bool need_feat = !(a.flags & DISTEGNN_FLAG_LAST) && a.g_aggv != nullptr;

    if (tid < H) {
        w1rs[tid] = a.w1r[tid];
        b2s[tid] = a.b2[tid];
        bxvs[tid] = a.bxv[tid];
        w3xvs[tid] = a.w3xv[tid];
        bxs[tid] = a.bx[tid];
        w3xs[tid] = a.w3x[tid];
    }
    for (int i = tid; i < 6 * H + VT_GROUPS * (VT_MAXC * H + 4 * VT_MAXC); i += VT_THREADS) gw1r[i] = 0.f;
    if (tid == 0) {
        for (int i = 0; i < 3 * VT_GROUPS; ++i) mbar_init(&bars[i], 1);
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
    const uint32_t lane_off = ((uint32_t)(32 * wq)) << 16;
    const uint32_t tA_hi = lane_off + col0, tA_lo = lane_off + col0 + 32, tD = lane_off + col0 + 64;
    const uint32_t tZ1 = lane_off + col0 + 128, tZ2 = lane_off + col0 + 192;
    __half* myslots = slots + grp * 2 * VT_IMG;
    float* Gt = tiles + grp * 2 * TILE_M * LDA;
    float* At = Gt + TILE_M * LDA;
    float* accG = accG_all + grp * VT_MAXC * H;
    float* accX = accX_all + grp * 4 * VT_MAXC;
    float* gdX = gdX_all + grp * TILE_M * 4;
    float* vrs = vrs_all + grp * TILE_M;
    int* sgraph = sgraph_all + grp * TILE_M;
    uint64_t* mbar = bars + 3 * grp;
    uint64_t* wbar = mbar + 1;                                   // [2]: weight slot filled
    const uint32_t bar_id = 1 + grp;
    uint32_t mph = 0;
    const uint32_t idesc = make_idesc_f16(128, 64, 0, 0);

    const int64_t num_tiles = (a.N + TN - 1) / TN;
    const int64_t tstride = (int64_t)gridDim.x * VT_GROUPS;
    const int64_t tile0 = (int64_t)blockIdx.x * VT_GROUPS + grp;
    const int64_t my_tiles = tile0 < num_tiles ? (num_tiles - tile0 + tstride - 1) / tstride : 0;
    const int64_t total_q = 6 * my_tiles;                        // matrices this group will consume, in order
    int64_t q_next = 0;                                          // next matrix to be requested (thread 0 of the group)
    auto request = [&]() {                                       // thread 0: stream matrix q_next into slot q_next & 1
        if (q_next < total_q) {
            const int slot = (int)(q_next & 1), m = (int)(q_next % 6);
            mbar_expect_tx(wbar + slot, VT_IMG * 2);
            bulk_g2s(myslots + slot * VT_IMG, a.wimg + (size_t)m * VT_IMG, VT_IMG * 2, wbar + slot);
            ++q_next;
        }
    };
    if (t == 0) {
        request();
        request();
    }
    int64_t q_use = 0;                                           // next matrix to be used (same on all threads)
    // publish the A operand, then thread 0 waits for the weight slot and issues the three split products
    auto issue = [&](bool accumulate) {
        wait_st();
        fence_before_sync();
        named_bar(bar_id, VT_GROUP);
        if (t == 0) {
            const int slot = (int)(q_use & 1);
            mbar_wait(wbar + slot, (uint32_t)((q_use >> 1) & 1));
            fence_after_sync();
            const __half* whi = myslots + slot * VT_IMG;
            tc16::issue_f16x3<VT_LBO>(col0 + 64u, col0, col0 + 32u, make_b_desc(smem_u32(whi), VT_LBO, 128),
                                      make_b_desc(smem_u32(whi + 64 * 64), VT_LBO, 128), idesc, accumulate);
            mma_commit(mbar);
        }
        __syncwarp();
        ++q_use;
    };
    auto mma_done = [&]() {                                      // ... and the slot it read is refilled two matrices ahead
        mbar_wait(mbar, mph);
        mph ^= 1;
        __syncwarp();
        fence_after_sync();
        if (t == 0) request();
    };
    auto colsum_G = [&](float* acc) {
        const int c = t & 63, h = t >> 6;
        float s0 = 0.f, s1 = 0.f;
        for (int e = 64 * h; e < 64 * h + 64; e += 2) {
            s0 += Gt[e * LDA + c];
            s1 += Gt[(e + 1) * LDA + c];
        }
        atomicAdd(acc + c, s0 + s1);
    };
    int cur_graph = -1;
    auto flush = [&](int g) {                                    // all threads of the group; caller synchronises
        if (g >= 0) {
            for (int i = t; i < C * H; i += VT_GROUP) {
                atomicAdd(a.g_G + (size_t)g * C * H + i, accG[i]);
                accG[i] = 0.f;
            }
            if (t < 3 * C) {
                const int d = t / C, c = t - d * C;
                atomicAdd(a.g_Xv + (size_t)g * 3 * C + t, accX[d * VT_MAXC + c]);
                accX[d * VT_MAXC + c] = 0.f;
            }
        }
    };

    float gW2[8][4], gWxv[8][4], gWx[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) gW2[i][j] = gWxv[i][j] = gWx[i][j] = 0.f;

    for (int64_t tile = tile0; tile < num_tiles; tile += tstride) {
        const int64_t n0 = tile * TN;
        const int nvalid = (int)min((int64_t)TN, a.N - n0);
        const int rows = nvalid * C;
        if (t < TN) sgraph[t] = (t < nvalid) ? __ldg(a.batch + n0 + t) : -1;
        named_bar(bar_id, VT_GROUP);
        const int g_first = sgraph[0];
        const bool single = (g_first == sgraph[nvalid - 1]);
        if (single && g_first != cur_graph) {
            flush(cur_graph);
            cur_graph = g_first;
            named_bar(bar_id, VT_GROUP);
        }

        // ---- the thread's row: node, channel, geometry, upstream scalars -----------------------------------------------
        const bool rvalid = t < rows;
        const int nl = rvalid ? t / C : 0;
        const int ch = rvalid ? t - nl * C : 0;
        const int g = rvalid ? sgraph[nl] : g_first;
        const size_t node = (size_t)(n0 + nl);
        float dx, dy, dz, vr, gpxv = 0.f, gpx = 0.f;
        float4 gt = make_float4(0.f, 0.f, 0.f, 0.f);
        float gv0 = 0.f, gv1 = 0.f, gv2 = 0.f;
        {
            const float4 xi = ldg4(a.x4 + node * 4);
            const float* Xg = a.Xv + (size_t)g * 3 * C;
            dx = __ldg(Xg + ch) - xi.x; dy = __ldg(Xg + C + ch) - xi.y; dz = __ldg(Xg + 2 * C + ch) - xi.z;
            vr = sqrtf(dx * dx + dy * dy + dz * dz);
            if (rvalid) {
                gt = ldg4(a.g_transv + node * 4);
                const float* gv = a.g_vsum + (size_t)g * K + 4;
                gv0 = __ldg(gv + ch); gv1 = __ldg(gv + C + ch); gv2 = __ldg(gv + 2 * C + ch);
                gpxv = -(gt.x * dx + gt.y * dy + gt.z * dz) * invC;
                gpx = gv0 * dx + gv1 * dy + gv2 * dz;
            }
            vrs[t] = rvalid ? vr : 0.f;
        }

        float v[64];
        // ---- stage 1: z1 -> TMEM; a1 = SiLU(z1) -> A;  MMA: z2 = a1·W2vᵀ ---------------------------------------------------
        {
            const float* hrow = a.Hn + node * H;
            const float* grow = a.G + ((size_t)g * C + ch) * H;
#pragma unroll
            for (int j4 = 0; j4 < 16; ++j4) {
                float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                if (rvalid) z = fma4(vr, *reinterpret_cast<const float4*>(w1rs + 4 * j4), add4(ldg4(hrow + 4 * j4), ldg4(grow + 4 * j4)));
                v[4 * j4] = z.x; v[4 * j4 + 1] = z.y; v[4 * j4 + 2] = z.z; v[4 * j4 + 3] = z.w;
            }
            tmem_store_row(tZ1, v);
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = rvalid ? silu(v[j]) : 0.f;
        }
        const float inv1 = encode_row_regs(v, tA_hi, tA_lo);
        issue(false);                                 // W2v
        mma_done();

        // ---- stage 2: z2 = D/s + b2v -> TMEM; mv = SiLU(z2) -> activation tile + A;  MMA: zxv = mv·Wxvᵀ ---------------------
        tmem_load_row(tD, v);
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] = fmaf(v[j], inv1, b2s[j]);
        tmem_store_row(tZ2, v);
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] = silu(v[j]);
        smem_store_row(At + t * LDA, v);
        const float inv2 = encode_row_regs(v, tA_hi, tA_lo);
        issue(false);                                 // Wxv
        mma_done();

        // ---- head xv: φ_xv, g_w3xv, g_zxv -> gradient tile (A keeps mv);  MMA: zx = mv·Wxᵀ ------------------------------------
        float phixv = 0.f, phix = 0.f;
        tmem_load_row(tD, v);
        {
            float u[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                const float zc = fmaf(v[j], inv2, bxvs[j]);
                const float s = sigmoid_f(zc);
                const float ac = zc * s, w3j = w3xvs[j];
                phixv = fmaf(ac, w3j, phixv);
                u[j] = gpxv * ac;
                v[j] = gpxv * w3j * (s * fmaf(zc, 1.0f - s, 1.0f));
            }
            warp_colsum64(u, lane);
            atomicAdd(gw3xv + 2 * lane, u[0]);
            atomicAdd(gw3xv + 2 * lane + 1, u[1]);
        }
        smem_store_row(Gt + t * LDA, v);
        float fm = row_absmax(v);                     // the two heads' gradient rows share one scale (one accumulator)
        issue(false);                                 // Wx (A unchanged: still mv); the barrier inside publishes both tiles
        wgrad128(gWxv, Gt, At, t);                    // g_Wxv += g_zxvᵀ·mv while the MMA runs
        colsum_G(gbxv);
        named_bar(bar_id, VT_GROUP);                  // gradient tile fully read
        mma_done();

        // ---- head x: φ_X, g_w3x, g_zx;  MMAs: g_mv = g_zxv·Wxv + g_zx·Wx -----------------------------------------------------
        float gzx[64];
        tmem_load_row(tD, gzx);
        {
            float u[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                const float zc = fmaf(gzx[j], inv2, bxs[j]);
                const float s = sigmoid_f(zc);
                const float ac = zc * s, w3j = w3xs[j];
                phix = fmaf(ac, w3j, phix);
                u[j] = gpx * ac;
                gzx[j] = gpx * w3j * (s * fmaf(zc, 1.0f - s, 1.0f));
            }
            warp_colsum64(u, lane);
            atomicAdd(gw3x + 2 * lane, u[0]);
            atomicAdd(gw3x + 2 * lane + 1, u[1]);
        }
        fm = row_absmax(gzx, fm);
        float sc, inv3;
        row_scale(fm, sc, inv3);
#pragma unroll
        for (int j4 = 0; j4 < 16; ++j4) {             // g_zxv back from the own row of the gradient tile (kept out of the
            const float4 q4 = *reinterpret_cast<const float4*>(Gt + t * LDA + 4 * j4);   // registers during the head)
            v[4 * j4] = q4.x; v[4 * j4 + 1] = q4.y; v[4 * j4 + 2] = q4.z; v[4 * j4 + 3] = q4.w;
        }
        encode_row_scaled(v, sc, tA_hi, tA_lo);       // A = g_zxv
        issue(false);                                 // Wxvᵀ
        smem_store_row(Gt + t * LDA, gzx);            // the gradient tile now holds g_zx (its readers passed the barrier above)
        mma_done();
        encode_row_scaled(gzx, sc, tA_hi, tA_lo);     // A = g_zx
        issue(true);                                  // Wxᵀ, accumulating; the barrier inside publishes the g_zx tile
        wgrad128(gWx, Gt, At, t);                     // g_Wx += g_zxᵀ·mv
        colsum_G(gbx);
        named_bar(bar_id, VT_GROUP);                  // both tiles fully read
        mma_done();

        // ---- g_z2 = (g_mv + upstream) ⊙ SiLU'(z2) -> gradient tile + A; a1 -> activation tile;  MMA: g_a1 = g_z2·W2v ------------
        float inv4;
        {
            tmem_load_row(tD, v);
            tmem_load_row(tZ2, gzx);                  // reuse as z2
            const bool up = need_feat && rvalid;
            const float* ga = a.g_aggv + node * H;
            const float* gs = a.g_vsum + (size_t)g * K + 4 + 3 * C + ch * H;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                float gm = v[j] * inv3;
                if (up) gm += fmaf(__ldg(ga + j), invC, __ldg(gs + j));
                v[j] = gm * dsilu(gzx[j]);
            }
            smem_store_row(Gt + t * LDA, v);
            inv4 = encode_row_regs(v, tA_hi, tA_lo);
            tmem_load_row(tZ1, gzx);                  // z1 -> a1 row for the weight gradient
#pragma unroll
            for (int j = 0; j < 64; ++j) gzx[j] = rvalid ? silu(gzx[j]) : 0.f;
            smem_store_row(At + t * LDA, gzx);
        }
        issue(false);                                 // W2vᵀ
        wgrad128(gW2, Gt, At, t);                     // g_W2v += g_z2ᵀ·a1
        colsum_G(gb2);
        named_bar(bar_id, VT_GROUP);
        mma_done();

        // ---- g_z1 = D/s ⊙ SiLU'(z1) -> gradient tile; g_vr; geometry gradient ---------------------------------------------------
        {
            tmem_load_row(tD, v);
            tmem_load_row(tZ1, gzx);
            float gr = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                v[j] = v[j] * inv4 * dsilu(gzx[j]);
                gr = fmaf(v[j], w1rs[j], gr);
            }
            fence_before_sync();                      // D reads ordered before the next tile's first MMA
            smem_store_row(Gt + t * LDA, v);
            // gΔX = −g_trans_v·φ_xv/C + g_vsum·φ_X + g_vr·ΔX/‖ΔX‖
            const float s1 = -phixv * invC, s3 = vr > 0.f ? gr / vr : 0.f;
            float4 gd = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rvalid) gd = make_float4(fmaf(gt.x, s1, fmaf(gv0, phix, s3 * dx)), fmaf(gt.y, s1, fmaf(gv1, phix, s3 * dy)),
                                         fmaf(gt.z, s1, fmaf(gv2, phix, s3 * dz)), 0.f);
            *reinterpret_cast<float4*>(gdX + 4 * t) = gd;
        }
        named_bar(bar_id, VT_GROUP);                  // g_z1 tile and gΔX visible

        // ---- reductions of the g_z1 tile and of gΔX ------------------------------------------------------------------------------
        {
            const int c64 = t & 63, h = t >> 6;
            for (int n = h; n < nvalid; n += 2) {           // g_Hn[node] = Σ_c g_z1
                float s = 0.f;
                for (int c = 0; c < C; ++c) s += Gt[(n * C + c) * LDA + c64];
                a.g_Hn[(size_t)(n0 + n) * H + c64] = s;
            }
            if (single) {                                    // Σ_i g_z1 per channel -> g_G
                for (int c = h; c < C; c += 2) {
                    float s = 0.f;
                    for (int n = 0; n < nvalid; ++n) s += Gt[(n * C + c) * LDA + c64];
                    accG[c * H + c64] += s;
                }
            } else {
                for (int n = h; n < nvalid; n += 2)
                    for (int c = 0; c < C; ++c)
                        atomicAdd(a.g_G + ((size_t)sgraph[n] * C + c) * H + c64, Gt[(n * C + c) * LDA + c64]);
            }
            float sr = 0.f;                                  // g_w_vr[n] += Σ_rows g_z1[row][n]·‖ΔX‖_row
            for (int q = 64 * h; q < 64 * h + 64; ++q) sr = fmaf(Gt[q * LDA + c64], vrs[q], sr);
            atomicAdd(gw1r + c64, sr);
        }
        if (t < nvalid) {                                    // g_x (virtual part) = −Σ_c gΔX
            float sx = 0.f, sy = 0.f, sz = 0.f;
            for (int c = 0; c < C; ++c) {
                const float4 gg = *reinterpret_cast<const float4*>(gdX + 4 * (t * C + c));
                sx += gg.x; sy += gg.y; sz += gg.z;
            }
            *reinterpret_cast<float4*>(a.g_xv + (size_t)(n0 + t) * 4) = make_float4(-sx, -sy, -sz, 0.f);
        }
        if (t >= 64 && t < 64 + 3 * C) {                     // g_Xv[b,d,c] += Σ_i gΔX_d
            const int k = t - 64, d = k / C, c = k - d * C;
            if (single) {
                float s = 0.f;
                for (int n = 0; n < nvalid; ++n) s += gdX[4 * (n * C + c) + d];
                accX[d * VT_MAXC + c] += s;
            } else {
                for (int n = 0; n < nvalid; ++n) atomicAdd(a.g_Xv + (size_t)sgraph[n] * 3 * C + k, gdX[4 * (n * C + c) + d]);
            }
        }
        named_bar(bar_id, VT_GROUP);                  // tiles and per-row arrays are rewritten by the next iteration
    }
    flush(cur_graph);

    wgrad128_flush(a.g_w2, gW2, t);
    wgrad128_flush(a.g_wxv, gWxv, t);
    wgrad128_flush(a.g_wx, gWx, t);
    fence_before_sync();
    __syncthreads();
    if (tid < H) {
        atomicAdd(a.g_w1r + tid, gw1r[tid]);
        atomicAdd(a.g_b2 + tid, gb2[tid]);
        atomicAdd(a.g_bxv + tid, gbxv[tid]);
        atomicAdd(a.g_w3xv + tid, gw3xv[tid]);
        atomicAdd(a.g_bx + tid, gbx[tid]);
        atomicAdd(a.g_w3x + tid, gw3x[tid]);
    }
    if ((tid >> 5) == 0) tmem_dealloc(tbase, 512);
}

}  // namespace degnn

extern "C" int distegnn_virtual_bwd_prepare(int A, int C, int Na, const float* layer_params, void* weight_images,
                                            void* stream) {
    using namespace degnn;
    if (int rc = check_dims(A, C, Na)) return rc;
    DEGNN_CHECK_ARG(layer_params && weight_images, "null pointer");
    Layout L = make_layout(A, C, Na);
    virtual_bwd_images_kernel<<<6, 256, 0, (cudaStream_t)stream>>>(layer_params + L.off[DISTEGNN_P_V_W2],
                                                                   layer_params + L.off[DISTEGNN_P_V_WXV],
                                                                   layer_params + L.off[DISTEGNN_P_V_WX],
                                                                   reinterpret_cast<__half*>(weight_images));
    DEGNN_CHECK_LAUNCH();
    return DISTEGNN_OK;
}

extern "C" int distegnn_virtual_layer_bwd(int64_t n_nodes, int n_graphs, int A, int C, int Na, unsigned flags,
                                          const int32_t* batch32, const float* x4, const float* Hn, const float* Xv,
                                          const float* G, const float* layer_params, const void* weight_images,
                                          const float* g_agg_v, const float* g_trans_v, const float* g_vsum,
                                          float* g_Hn, float* g_xv, float* g_G, float* g_Xv, float* g_layer_params,
                                          void* stream) {
    using namespace degnn;
    if (int rc = check_dims(A, C, Na)) return rc;
    if (n_nodes == 0) return DISTEGNN_OK;
    DEGNN_CHECK_ARG(n_nodes > 0 && n_graphs > 0, "bad size");
    DEGNN_CHECK_ARG(batch32 && x4 && Hn && Xv && G && layer_params && weight_images && g_trans_v && g_vsum && g_Hn && g_xv &&
                        g_G && g_Xv && g_layer_params,
                    "null pointer");
    Layout L = make_layout(A, C, Na);
    VirtBwdTcArgs a;
    a.N = n_nodes; a.B = n_graphs; a.C = C; a.flags = flags;
    a.batch = batch32; a.x4 = x4; a.Hn = Hn; a.Xv = Xv; a.G = G;
    a.w1r = layer_params + L.off[DISTEGNN_P_V_W1R];
    a.b2 = layer_params + L.off[DISTEGNN_P_V_B2];
    a.bxv = layer_params + L.off[DISTEGNN_P_V_BXV];
    a.w3xv = layer_params + L.off[DISTEGNN_P_V_W3XV];
    a.bx = layer_params + L.off[DISTEGNN_P_V_BX];
    a.w3x = layer_params + L.off[DISTEGNN_P_V_W3X];
    a.wimg = reinterpret_cast<const __half*>(weight_images);
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
    ensure_dynamic_smem((const void*)virtual_layer_bwd_tc_kernel, (int)VT_SMEM_BYTES);
    const int TN = TILE_M / C;
    const int64_t tiles = (n_nodes + TN - 1) / TN;
    int64_t grid = (tiles + VT_GROUPS - 1) / VT_GROUPS;
    if (grid > sm_count()) grid = sm_count();
    virtual_layer_bwd_tc_kernel<<<(unsigned)grid, VT_THREADS, VT_SMEM_BYTES, (cudaStream_t)stream>>>(a);
    DEGNN_CHECK_LAUNCH();
    return DISTEGNN_OK;
}
