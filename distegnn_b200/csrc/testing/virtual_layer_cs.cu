// Real<->virtual stage, column-split flavour — behind distegnn_virtual_layer_fwd_cs (twin / experiment).
// Same math, layout and outputs as virtual_layer_tc16.cu (the production kernel behind distegnn_virtual_layer_fwd);
// replaces reference models/FastEGNN.py:252-253 (virtual geometry), 154-163 (edge_mode_virtual), 180, 191-193, 207,
// 220-223 (virtual halves of coord_model_vel / coord_model_virtual / node_model / node_model_virtual) and the
// global_mean_pool scatters at :193,:222.
//
// Result first: 1.71 ms vs 1.68 ms for the thread-per-row kernel at config 5 — issue slots 43 % -> 54 % busy but 29 %
// more instructions (per-thread overheads double), so it is NOT the production kernel.  The hypothesis was: the
// thread-per-row kernel holds 16 warps per SM (128 registers) and ncu shows its warps issuing only every
// ~9 cycles (five group barriers and three MMA round trips per 16-node tile, first-layer loads, MUFU) — issue slots are
// 43 % busy.  Here two threads share a row (node, channel), each owning 32 of its 64 columns, at 64 registers: 1024
// threads = 32 warps per SM with the same TMEM (4 tile groups x 128 columns) and shared-memory footprint.  The scheme is
// the one of edge_layer_cs.cu: warp k of an 8-warp group has TMEM lane quarter k & 3 and column half k >> 2; the per-row
// range scale is agreed through flags posted before the group barrier (cold path: row maxima exchanged through shared
// memory); the two half-row partial sums of the 1-wide heads are combined through shared memory at the barrier that
// already follows each head.
//   stage 1  a1 = SiLU(Hn[node] + G[graph,c] + w_r·‖ΔX‖)  -> A              MMA 1: D = a1·W2vᵀ
//   stage 2  mv = SiLU(D + b2v) -> shared tile + A                            MMA 2: D = mv·Wxvᵀ
//            (while it runs: agg_v[node] = mean_c mv, per-graph Σ_i mv accumulated in shared memory)
//   stage 3a φ_xv = w3xv·SiLU(D + bxv)                                        MMA 3: D = mv·WXᵀ
//   stage 3b φ_X  = w3x·SiLU(D + bx);  trans_v[node] = mean_c(−ΔX·φ_xv);  per-graph Σ_i ΔX·φ_X accumulated.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc16.cuh"
#include "umma.cuh"

namespace degnn {

struct VirtCsArgs {
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

constexpr int VC_THREADS = 1024, VC_GROUPS = 4, VC_GROUP = 256, VC_WARPS = 8;
constexpr int VC_ROW = 68;
constexpr int VC_MAXC = DISTEGNN_MAX_CHANNELS;
constexpr int VC_W = 4096;
constexpr int VC_SMEM_BYTES = 6 * VC_W * 2                            // W2v, Wxv, WX (hi+lo)
                              + VC_GROUPS * TILE_M * VC_ROW * 4       // mv tile per group
                              + VC_GROUPS * VC_MAXC * H * 4           // Σ mv accumulators per group
                              + VC_GROUPS * 4 * VC_MAXC * 4           // Σ ΔX·φ_X accumulators per group
                              + 6 * H * 4                             // w1r, b2v, bxv, w3xv, bx, w3x
                              + VC_GROUPS * TILE_M * 4 * 4            // ΔX per row
                              + VC_GROUPS * 4 * TILE_M * 4            // φ_xv, φ_X per row and column half
                              + VC_GROUPS * TILE_M * 4                // graph id per local node
                              + VC_GROUPS * 2 * VC_WARPS * 4          // out-of-range flags per warp, one set per stage
                              + VC_GROUPS * 2 * TILE_M * 4            // row maxima of the two column halves (cold path)
                              + 128;
constexpr uint32_t VC_LBO = 1024;
using tc16::kFast;
using tc16::kSafe;

// 8 fp32 values held as 4 register pairs (·s) -> 4 packed hi words + 4 packed lo words; mx tracks max |hi|
template <bool SCALED>
__device__ __forceinline__ void vsplit8(const f32x2 (&v)[4], float s, uint32_t (&hi)[4], uint32_t (&lo)[4], __half2& mx) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const f32x2 x = SCALED ? mul2(v[j], bc2(s)) : v[j];
        float x0, x1, l0, l1;
        upk2(x, x0, x1);
        const __half2 h = __floats2half2_rn(x0, x1);
        const float2 hf = __half22float2(h);
        upk2(sub2(x, pk2(hf.x, hf.y)), l0, l1);
        const __half2 l = __floats2half2_rn(l0, l1);
        mx = __hmax2(mx, __habs2(h));
        hi[j] = *reinterpret_cast<const uint32_t*>(&h);
        lo[j] = *reinterpret_cast<const uint32_t*>(&l);
    }
}
__device__ __forceinline__ float vmax8(const f32x2 (&v)[4], float fm) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float v0, v1;
        upk2(v[j], v0, v1);
        fm = fmaxf(fm, fmaxf(fabsf(v0), fabsf(v1)));
    }
    return fm;
}

__global__ void __launch_bounds__(VC_THREADS, 1) virtual_layer_cs_kernel(const VirtCsArgs a) {
    using namespace umma;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __half* W2hi = reinterpret_cast<__half*>(smem_raw);
    __half* W2lo = W2hi + VC_W;
    __half* Wxvhi = W2lo + VC_W;
    __half* Wxvlo = Wxvhi + VC_W;
    __half* Wxhi = Wxvlo + VC_W;
    __half* Wxlo = Wxhi + VC_W;
    float* tiles = reinterpret_cast<float*>(Wxlo + VC_W);
    float* accH_all = tiles + VC_GROUPS * TILE_M * VC_ROW;
    float* accX_all = accH_all + VC_GROUPS * VC_MAXC * H;
    float* w1rs = accX_all + VC_GROUPS * 4 * VC_MAXC;
    float* b2s = w1rs + H;
    float* bxvs = b2s + H;
    float* w3xvs = bxvs + H;
    float* bxs = w3xvs + H;
    float* w3xs = bxs + H;
    float* dX_all = w3xs + H;
    float* phi_all = dX_all + VC_GROUPS * TILE_M * 4;                 // [4][xv: half0 128 | half1 128 | x: half0 | half1]
    int* sgraph_all = reinterpret_cast<int*>(phi_all + VC_GROUPS * 4 * TILE_M);
    uint32_t* oflag_all = reinterpret_cast<uint32_t*>(sgraph_all + VC_GROUPS * TILE_M);   // [4][2][8]
    float* rowmax_all = reinterpret_cast<float*>(oflag_all + VC_GROUPS * 2 * VC_WARPS);    // [4][2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(rowmax_all + VC_GROUPS * 2 * TILE_M);
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + VC_GROUPS);

    const int tid = threadIdx.x;
    const int grp = tid >> 8, tg = tid & 255, wk = tg >> 5, lane = tid & 31;
    const int wq = wk & 3, hf = wk >> 2;
    const int r = 32 * wq + lane;              // row (node, channel) of the tile shared with the thread of the other half
    const int cb = 32 * hf;                    // first owned column
    const int C = a.C;
    const int K = 4 + 3 * C + H * C;
    const bool need_feat = !(a.flags & DISTEGNN_FLAG_LAST);
    const int TN = TILE_M / C;

    // ---- one-time setup ---------------------------------------------------------------------------
    tc16::stage_weight<VC_THREADS>(W2hi, W2lo, a.w2, 0, 64, tid);
    tc16::stage_weight<VC_THREADS>(Wxvhi, Wxvlo, a.wxv, 0, 64, tid);
    tc16::stage_weight<VC_THREADS>(Wxhi, Wxlo, a.wx, 0, 64, tid);
    if (tid < H) {
        w1rs[tid] = a.w1r[tid];
        b2s[tid] = a.b2[tid];
        bxvs[tid] = a.bxv[tid];
        w3xvs[tid] = a.w3xv[tid];
        bxs[tid] = a.bx[tid];
        w3xs[tid] = a.w3x[tid];
    }
    for (int i = tid; i < VC_GROUPS * (VC_MAXC * H + 4 * VC_MAXC); i += VC_THREADS) accH_all[i] = 0.f;
    if (tid == 0) {
        for (int i = 0; i < VC_GROUPS; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    __syncwarp();
    if ((tid >> 5) == 0) tmem_alloc(tmem_base_s, 512);
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();

    const uint32_t tbase = *tmem_base_s;
    const uint32_t lane_off = ((uint32_t)(32 * wq)) << 16;
    const uint32_t col0 = tbase + (uint32_t)grp * 128u;
    const uint32_t tA_hi = lane_off + col0 + 16u * hf, tA_lo = lane_off + col0 + 32u + 16u * hf;
    const uint32_t tD = lane_off + col0 + 64u + 32u * hf;
    float* tile_s = tiles + grp * TILE_M * VC_ROW;
    float* myrow = tile_s + r * VC_ROW + cb;
    float* accH = accH_all + grp * VC_MAXC * H;
    float* accX = accX_all + grp * 4 * VC_MAXC;
    float* dXs = dX_all + grp * TILE_M * 4;
    float* phis = phi_all + grp * 4 * TILE_M;
    int* sgraph = sgraph_all + grp * TILE_M;
    uint32_t* oflag = oflag_all + grp * 2 * VC_WARPS;
    float* rowmax = rowmax_all + grp * 2 * TILE_M;
    uint64_t* mbar = bars + grp;
    const uint32_t bar_id = 1 + grp;
    uint32_t mph = 0;
    int cur_graph = -1;

    auto flush = [&](int g) {                      // all threads of the group
        if (g >= 0) {
            float* dst = a.vsum + (size_t)g * K;
            if (need_feat)
                for (int i = tg; i < C * H; i += VC_GROUP) {
                    atomicAdd(dst + 4 + 3 * C + i, accH[i]);
                    accH[i] = 0.f;
                }
            if (tg < 3 * C) {
                atomicAdd(dst + 4 + tg, accX[tg]);
                accX[tg] = 0.f;
            }
        }
    };
    auto issue = [&](const __half* whi, const __half* wlo) {
        if (tg == 0) {
            fence_after_sync();
            const uint32_t idesc = make_idesc_f16(128, 64, 0, 0);
            tc16::issue_f16x3<VC_LBO>(col0 + 64u, col0, col0 + 32u, make_b_desc(smem_u32(whi), VC_LBO, 128),
                                      make_b_desc(smem_u32(wlo), VC_LBO, 128), idesc, false);
            mma_commit(mbar);
        }
        __syncwarp();
    };
    auto mma_done = [&]() {
        mbar_wait(mbar, mph);
        mph ^= 1;
        __syncwarp();
        fence_after_sync();
    };
    auto group_flag = [&](int stage) {
        const uint4 f0 = *reinterpret_cast<const uint4*>(oflag + stage * VC_WARPS);
        const uint4 f1 = *reinterpret_cast<const uint4*>(oflag + stage * VC_WARPS + 4);
        return ((f0.x | f0.y | f0.z | f0.w) | (f1.x | f1.y | f1.z | f1.w)) != 0u;
    };
    // Encode the own half row, produced 8 columns at a time by f(j, v, first_pass, flavour, qmax), into the A operand.
    // Hot pass with the batched-reciprocal SiLU; every warp posts "out of range" before the group barrier `bar`; if any
    // warp of the group did, all redo their rows with the per-element flavour and ONE power-of-two scale per row
    // (maxima of the two halves exchanged through shared memory).  Returns 1/scale.  Ends after a group barrier.
    auto encode_half = [&](auto&& f, int stage) -> float {
        float qmax = 0.f;
        __half2 mx = __floats2half2_rn(0.f, 0.f);
#pragma unroll 2
        for (int j = 0; j < 4; ++j) {
            f32x2 v[4];
            uint32_t hi[4], lo[4];
            f(j, v, true, kFast, qmax);
            vsplit8<false>(v, 1.0f, hi, lo, mx);
            tmem_st4(tA_hi + 4 * j, hi);
            tmem_st4(tA_lo + 4 * j, lo);
        }
        const bool bad = __any_sync(FULL, tc16::row_overflow(mx) || silu_q_overflow(qmax));
        if (lane == 0) oflag[stage * VC_WARPS + wk] = bad ? 1u : 0u;
        wait_st();
        fence_before_sync();
        named_bar(bar_id, VC_GROUP);
        float inv = 1.0f;
        if (group_flag(stage)) {                   // cold
            float fm = 0.f, sc;
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                f32x2 v[4];
                f(j, v, true, kSafe, qmax);        // first_pass again: side effects are rewritten with valid values
                fm = vmax8(v, fm);
            }
            rowmax[hf * TILE_M + r] = fm;
            named_bar(bar_id, VC_GROUP);
            tc16::range_scale(fmaxf(rowmax[r], rowmax[TILE_M + r]), sc, inv);
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                f32x2 v[4];
                uint32_t hi[4], lo[4];
                f(j, v, false, kSafe, qmax);
                vsplit8<true>(v, sc, hi, lo, mx);
                tmem_st4(tA_hi + 4 * j, hi);
                tmem_st4(tA_lo + 4 * j, lo);
            }
            wait_st();
            fence_before_sync();
            named_bar(bar_id, VC_GROUP);
        }
        return inv;
    };

    const int64_t num_tiles = (a.N + TN - 1) / TN;
    const int64_t tstride = (int64_t)gridDim.x * VC_GROUPS;
    for (int64_t tile = (int64_t)blockIdx.x * VC_GROUPS + grp; tile < num_tiles; tile += tstride) {
        const int64_t n0 = tile * TN;
        const int nvalid = (int)min((int64_t)TN, a.N - n0);
        const int rows = nvalid * C;
        {   // pull the next tile's inputs (a contiguous block of Hn rows, x4, graph ids) into L1 while this one computes
            const int64_t nn0 = (tile + tstride) * TN;
            const int nnv = (int)min((int64_t)TN, a.N - nn0);
            for (int i = tg; i < 2 * nnv; i += VC_GROUP) prefetch_l1(a.Hn + (size_t)nn0 * H + 32 * i);
            if (nnv > 0) {
                if (tg < (nnv * 16 + 127) / 128) prefetch_l1(a.x4 + (size_t)nn0 * 4 + 32 * tg);
                if (tg == 255) prefetch_l1(a.batch + nn0);
                if (tg == 254) prefetch_l1(a.batch + nn0 + nnv - 1);
            }
        }
        if (tg < TN) sgraph[tg] = (tg < nvalid) ? __ldg(a.batch + n0 + tg) : -1;
        named_bar(bar_id, VC_GROUP);
        const int g_first = sgraph[0], g_last = sgraph[nvalid - 1];
        const bool single = (g_first == g_last);
        if (single && g_first != cur_graph) {
            flush(cur_graph);
            cur_graph = g_first;
            named_bar(bar_id, VC_GROUP);
        }

        // ---- stage 1 ------------------------------------------------------------------------------
        const bool rvalid = r < rows;
        const int nl = rvalid ? r / C : 0;
        const int ch = rvalid ? r - nl * C : 0;
        const int g = rvalid ? sgraph[nl] : g_first;
        const size_t node = (size_t)(n0 + nl);
        float vr;
        {
            const float4 xi = ldg4(a.x4 + node * 4);
            const float* Xg = a.Xv + (size_t)g * 3 * C;
            const float dx = __ldg(Xg + ch) - xi.x, dy = __ldg(Xg + C + ch) - xi.y, dz = __ldg(Xg + 2 * C + ch) - xi.z;
            vr = sqrtf(dx * dx + dy * dy + dz * dz);
            if (hf == 0) *reinterpret_cast<float4*>(dXs + 4 * r) = make_float4(dx, dy, dz, 0.f);
        }
        const float* hrow = a.Hn + node * H + cb;
        const float* grow = a.G + ((size_t)g * C + ch) * H + cb;
        const f32x2 vr2 = bc2(vr);
        const float inv1 = encode_half(
            [&](int j, f32x2 (&v)[4], bool, auto safe, float& qmax) {
#pragma unroll
                for (int j4 = 0; j4 < 2; ++j4) {
                    const int cc = 8 * j + 4 * j4;
                    const ulonglong2 hh = __ldg(reinterpret_cast<const ulonglong2*>(hrow + cc));
                    const ulonglong2 gg = __ldg(reinterpret_cast<const ulonglong2*>(grow + cc));
                    const ulonglong2 wr = *reinterpret_cast<const ulonglong2*>(w1rs + cb + cc);
                    f32x2 p0 = fma2(vr2, wr.x, add2(hh.x, gg.x)), p1 = fma2(vr2, wr.y, add2(hh.y, gg.y));
                    silu4p<decltype(safe)::value>(p0, p1, qmax);
                    if (!rvalid) p0 = p1 = 0ull;
                    v[2 * j4] = p0;
                    v[2 * j4 + 1] = p1;
                }
            },
            0);
        issue(W2hi, W2lo);
        mma_done();

        // ---- stage 2: mv = SiLU(D + b2v) -> shared tile and A ---------------------------------------------
        const float inv2 = encode_half(
            [&](int j, f32x2 (&v)[4], bool first, auto safe, float& qmax) {
                uint32_t d[8];
                tmem_ld8(tD + 8 * j, d);
                wait_ld();
                const f32x2 is2 = bc2(inv1);
#pragma unroll
                for (int j4 = 0; j4 < 2; ++j4) {
                    const int cc = 8 * j + 4 * j4;
                    const ulonglong2 bb = *reinterpret_cast<const ulonglong2*>(b2s + cb + cc);
                    f32x2 m0 = fma2(pk2u(d[4 * j4 + 0], d[4 * j4 + 1]), is2, bb.x);
                    f32x2 m1 = fma2(pk2u(d[4 * j4 + 2], d[4 * j4 + 3]), is2, bb.y);
                    silu4p<decltype(safe)::value>(m0, m1, qmax);
                    if (first && need_feat) *reinterpret_cast<ulonglong2*>(myrow + cc) = make_ulonglong2(m0, m1);
                    v[2 * j4] = m0;
                    v[2 * j4 + 1] = m1;
                }
            },
            1);
        issue(Wxvhi, Wxvlo);
        // pools of mv while MMA 2 runs
        if (need_feat) {
            // thread <-> (column pair, eighth): one LDS.64 + one FADD2 per two elements, two chains per sum
            const int c2 = 2 * lane, q8 = wk;
            auto ld2 = [](const float* p) { return *reinterpret_cast<const f32x2*>(p); };
            const f32x2 invC2 = bc2(1.0f / (float)C);
            for (int n = q8; n < nvalid; n += VC_WARPS) {      // mean over channels per node
                const float* base = tile_s + (n * C) * VC_ROW + c2;
                f32x2 s0 = 0ull, s1 = 0ull;
                int c = 0;
                for (; c + 1 < C; c += 2) {
                    s0 = add2(s0, ld2(base + c * VC_ROW));
                    s1 = add2(s1, ld2(base + (c + 1) * VC_ROW));
                }
                if (c < C) s0 = add2(s0, ld2(base + c * VC_ROW));
                *reinterpret_cast<f32x2*>(a.agg_v + (size_t)(n0 + n) * H + c2) = mul2(add2(s0, s1), invC2);
            }
            if (single) {                                      // sum over nodes per channel
                for (int c = q8; c < C; c += VC_WARPS) {
                    const float* base = tile_s + c * VC_ROW + c2;
                    const int nstep = C * VC_ROW;
                    f32x2 s0 = 0ull, s1 = 0ull;
                    int n = 0;
                    for (; n + 1 < nvalid; n += 2) {
                        s0 = add2(s0, ld2(base + n * nstep));
                        s1 = add2(s1, ld2(base + (n + 1) * nstep));
                    }
                    if (n < nvalid) s0 = add2(s0, ld2(base + n * nstep));
                    f32x2* acc = reinterpret_cast<f32x2*>(accH + c * H + c2);
                    *acc = add2(*acc, add2(s0, s1));
                }
            } else {
                for (int n = q8; n < nvalid; n += VC_WARPS) {
                    float* dst = a.vsum + (size_t)sgraph[n] * K + 4 + 3 * C + c2;
                    for (int c = 0; c < C; ++c) {
                        float v0, v1;
                        upk2(ld2(tile_s + (n * C + c) * VC_ROW + c2), v0, v1);
                        atomicAdd(dst + c * H, v0);
                        atomicAdd(dst + c * H + 1, v1);
                    }
                }
            }
        }
        mma_done();

        // ---- the two 1-wide heads: each thread sums its 32 columns, the halves meet in shared memory ----------------
        auto head_pass = [&](const float* bs, const float* ws, auto safe, float& qmax) {
            f32x2 ph01 = bc2(0.f), ph23 = bc2(0.f);
            const f32x2 is2 = bc2(inv2);
#pragma unroll 2
            for (int j = 0; j < 4; ++j) {
                uint32_t d[8];
                tmem_ld8(tD + 8 * j, d);
                wait_ld();
#pragma unroll
                for (int j4 = 0; j4 < 2; ++j4) {
                    const int cc = cb + 8 * j + 4 * j4;
                    const ulonglong2 bb = *reinterpret_cast<const ulonglong2*>(bs + cc);
                    const ulonglong2 ww = *reinterpret_cast<const ulonglong2*>(ws + cc);
                    f32x2 s0 = fma2(pk2u(d[4 * j4 + 0], d[4 * j4 + 1]), is2, bb.x);
                    f32x2 s1 = fma2(pk2u(d[4 * j4 + 2], d[4 * j4 + 3]), is2, bb.y);
                    silu4p<decltype(safe)::value>(s0, s1, qmax);
                    ph01 = fma2(s0, ww.x, ph01);
                    ph23 = fma2(s1, ww.y, ph23);
                }
            }
            float p0, p1, p2, p3;
            upk2(ph01, p0, p1);
            upk2(ph23, p2, p3);
            return (p0 + p1) + (p2 + p3);
        };
        auto head = [&](const float* bs, const float* ws) {
            float qmax = 0.f;
            float phi = head_pass(bs, ws, kFast, qmax);
            if (kSiluGuard && __any_sync(FULL, silu_q_overflow(qmax))) phi = head_pass(bs, ws, kSafe, qmax);   // cold
            return phi;
        };
        // ---- stage 3a: φ_xv = w3xv·SiLU(D + bxv) --------------------------------------------------------
        phis[hf * TILE_M + r] = head(bxvs, w3xvs);
        fence_before_sync();
        named_bar(bar_id, VC_GROUP);                  // D fully read (A still holds mv); both halves of φ_xv visible
        issue(Wxhi, Wxlo);
        // trans_v[node] = mean_c(−ΔX_c·φ_xv,c) while MMA 3 runs
        for (int i = tg; i < nvalid * 3; i += VC_GROUP) {
            const int n = i / 3, d = i - 3 * n;
            float s = 0.f;
            for (int c = 0; c < C; ++c) {
                const int rr = n * C + c;
                s = fmaf(-dXs[4 * rr + d], phis[rr] + phis[TILE_M + rr], s);
            }
            a.trans_v[(size_t)(n0 + n) * 4 + d] = s / (float)C;
        }
        mma_done();

        // ---- stage 3b: φ_X = w3x·SiLU(D + bx); Σ_i ΔX_ic·φ_X,ic per graph [3][C] -----------------------------
        phis[(2 + hf) * TILE_M + r] = head(bxs, w3xs);
        fence_before_sync();
        named_bar(bar_id, VC_GROUP);
        if (tg < 3 * C) {
            const int d = tg / C, c = tg - d * C;
            const float* phx = phis + 2 * TILE_M;
            if (single) {
                float s = 0.f;
                for (int n = 0; n < nvalid; ++n) {
                    const int rr = n * C + c;
                    s = fmaf(dXs[4 * rr + d], phx[rr] + phx[TILE_M + rr], s);
                }
                accX[tg] += s;
            } else {
                for (int n = 0; n < nvalid; ++n) {
                    const int rr = n * C + c;
                    atomicAdd(a.vsum + (size_t)sgraph[n] * K + 4 + tg, dXs[4 * rr + d] * (phx[rr] + phx[TILE_M + rr]));
                }
            }
        }
        named_bar(bar_id, VC_GROUP);                  // sgraph/dXs/phis/tile_s are rewritten by the next tile
    }
    flush(cur_graph);

    fence_before_sync();
    __syncthreads();
    if ((tid >> 5) == 0) tmem_dealloc(tbase, 512);
}

}  // namespace degnn

extern "C" int distegnn_virtual_layer_fwd_cs(int64_t n_nodes, int n_graphs, int A, int C, int Na, unsigned flags,
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
    VirtCsArgs a;
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
    ensure_dynamic_smem((const void*)virtual_layer_cs_kernel, (int)VC_SMEM_BYTES);
    const int TN = TILE_M / C;
    const int64_t tiles = (n_nodes + TN - 1) / TN;
    int64_t grid = (tiles + VC_GROUPS - 1) / VC_GROUPS;
    if (grid > sm_count()) grid = sm_count();
    virtual_layer_cs_kernel<<<(unsigned)grid, VC_THREADS, VC_SMEM_BYTES, (cudaStream_t)stream>>>(a);
    DEGNN_CHECK_LAUNCH();
    return DISTEGNN_OK;
}
