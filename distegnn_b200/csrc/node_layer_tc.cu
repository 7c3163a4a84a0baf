// Node update of one E_GCL_vel layer on the tensor cores — production kernel behind distegnn_node_layer_fwd.
// Replaces reference models/FastEGNN.py:177-183 (sum of the three coordinate terms, φ_v) and :203-217
// (node_model) and emits the per-node operands of the NEXT layer's fused stages (P, Q, Hn; SURVEY §7 "W1 split").
// Same math/outputs as node_layer.cu's fp32-FMA kernel (kept as ..._simt for cross-checks).
//
// One CTA per SM, 256 threads = 2 tile groups of 4 warps; thread r of a group owns node r of the group's current
// 128-node tile (TMEM lane r).  All eight 64x64 layers run as kind::f16 tile GEMMs with the fp16 2-term split
// (tc16.cuh); weights (hi/lo, 128 KB) stay resident in shared memory.  The 256-byte rows of h, agg_m, agg_v are
// fetched by TMA bulk copies into a padded staging buffer, one array at a time (the kernel has few tiles per SM
// and is bound by HBM, not by these round trips).  TMEM columns per group (256):
//   A_hi [0,32)  A_lo [32,64)  D1 [64,128)  D2 [128,192)  D3 / fp32 copy of h [192,256)
// Sequence per tile:   h -> A, copy;  MMA: D1 = h·Lᵀ (φ_v), D2 = h·N1aᵀ
//                      agg_m/deg -> A; MMA: D2 += ·N1bᵀ      agg_v -> A; MMA: D2 += ·N1cᵀ
//                      t1 = SiLU(D2 + attr·N1d + b1) -> A;   MMA: D1 = t1·N2ᵀ
//                      h' = h + D1 + b2 -> HBM, -> A;        MMA: [D1|D2|D3] = h'·[W1a';W1b';W1vh']ᵀ (N = 192)
//                      P = D1 + b1', Q = D2, Hn = D3 -> HBM.   x' and Σ(x',1) are computed while MMA 1 runs.
#include "common.cuh"
#include "tc16.cuh"
#include "umma.cuh"

namespace degnn {

struct NodeTcArgs {
    int64_t N;
    int B, Na, K;
    unsigned flags;
    const int32_t* rowptr; const int32_t* batch;
    const float* h; const float* x4; const float* vel; const float* attr;
    const float* agg_m; const float* agg_x; const float* agg_v; const float* trans_v;
    const float* lw; const float* lb; const float* lw3; const float* lb3;          // φ_v
    const float* n1; const float* nb1; const float* n2; const float* nb2;          // node MLP
    const float* nw1a; const float* nxb1; const float* nw1b; const float* nw1h;    // next layer
    float* h_out; float* x4_out; float* P; float* Q; float* Hn; float* loc_out; float* vsum;
};

constexpr int NT_THREADS = 256, NT_GROUPS = 2, NT_GROUP = 128;
constexpr int NT_ROW = 68;
constexpr int NT_W = 4096;                                    // halfs per 64x64 matrix part
constexpr int NT_SMEM_BYTES = (2 * NT_W + 6 * NT_W + 2 * NT_W + 2 * 3 * NT_W) * 2   // L, N1a-c, N2, NEXT (hi+lo)
                              + NT_GROUPS * TILE_M * NT_ROW * 4                      // staging rows
                              + (5 * H + DISTEGNN_MAX_NODE_ATTR * H) * 4             // lb, lw3, nb1, nb2, nxb1, N1d
                              + NT_GROUPS * 8 * 4                                    // accS[4] + sg[2] (+pad) per group
                              + 128                                                  // mbarriers + tmem base
                              + 256;                                                 // a zero row (FLAG_ZERO_AGG)
constexpr uint32_t NT_LBO64 = 1024, NT_LBO192 = 3072;

__global__ void __launch_bounds__(NT_THREADS, 1) node_layer_tc_kernel(const NodeTcArgs a) {
    using namespace umma;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __half* Lhi = reinterpret_cast<__half*>(smem_raw);
    __half* Llo = Lhi + NT_W;
    __half* N1hi = Llo + NT_W;            // [3][4096]
    __half* N1lo = N1hi + 3 * NT_W;
    __half* N2hi = N1lo + 3 * NT_W;
    __half* N2lo = N2hi + NT_W;
    __half* NXhi = N2lo + NT_W;           // 192 x 64
    __half* NXlo = NXhi + 3 * NT_W;
    float* stage_all = reinterpret_cast<float*>(NXlo + 3 * NT_W);
    float* lbs = stage_all + NT_GROUPS * TILE_M * NT_ROW;
    float* lw3s = lbs + H;
    float* nb1s = lw3s + H;
    float* nb2s = nb1s + H;
    float* nxb1s = nb2s + H;
    float* n1ds = nxb1s + H;              // [Na][64]
    float* acc_all = n1ds + DISTEGNN_MAX_NODE_ATTR * H;          // per group: accS[4], sg[2] (as int), pad[2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(acc_all + NT_GROUPS * 8);   // [2 groups][2]: staging, mma
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 2 * NT_GROUPS);
    float* zero_row = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 128);   // 256 B of zeros, 16-byte aligned

    const int tid = threadIdx.x;
    const int grp = tid >> 7, t = tid & 127, lane = tid & 31, wq = (tid >> 5) & 3;
    const bool last = a.flags & DISTEGNN_FLAG_LAST;
    const bool zero_agg = a.flags & DISTEGNN_FLAG_ZERO_AGG;   // leave agg_m / agg_x zeroed for the next edge stage
    const int Na = a.Na;

    // ---- one-time setup ---------------------------------------------------------------------------
    tc16::stage_weight<NT_THREADS>(Lhi, Llo, a.lw, 0, 64, tid);
    if (!last) {
        for (int c = 0; c < 3; ++c) tc16::stage_weight<NT_THREADS>(N1hi + c * NT_W, N1lo + c * NT_W, a.n1 + c * H * H, 0, 64, tid);
        tc16::stage_weight<NT_THREADS>(N2hi, N2lo, a.n2, 0, 64, tid);
        tc16::stage_weight<NT_THREADS>(NXhi, NXlo, a.nw1a, 0, 192, tid);
        tc16::stage_weight<NT_THREADS>(NXhi, NXlo, a.nw1b, 64, 192, tid);
        tc16::stage_weight<NT_THREADS>(NXhi, NXlo, a.nw1h, 128, 192, tid);
    }
    if (tid < H) {
        lbs[tid] = a.lb[tid];
        lw3s[tid] = a.lw3[tid];
        if (!last) {
            nb1s[tid] = a.nb1[tid];
            nb2s[tid] = a.nb2[tid];
            nxb1s[tid] = a.nxb1[tid];
        }
    }
    if (!last)
        for (int i = tid; i < Na * H; i += NT_THREADS) n1ds[i] = a.n1[(size_t)3 * H * H + i];
    if (tid < NT_GROUPS * 8) acc_all[tid] = 0.f;
    if (tid < H) zero_row[tid] = 0.f;
    if (tid == 0) {
        for (int i = 0; i < 2 * NT_GROUPS; ++i) mbar_init(&bars[i], 1);
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
    const uint32_t col0 = tbase + (uint32_t)grp * 256u;
    const uint32_t tA_hi = col0, tA_lo = col0 + 32, tD1 = col0 + 64, tD2 = col0 + 128, tD3 = col0 + 192;
    const uint32_t idesc64 = make_idesc_f16(128, 64, 0, 0), idesc192 = make_idesc_f16(128, 192, 0, 0);
    auto desc = [&](const __half* p, uint32_t lbo) { return make_b_desc(smem_u32(p), lbo, 128); };
    const uint64_t dLhi = desc(Lhi, NT_LBO64), dLlo = desc(Llo, NT_LBO64);
    const uint64_t dN2hi = desc(N2hi, NT_LBO64), dN2lo = desc(N2lo, NT_LBO64);
    const uint64_t dNXhi = desc(NXhi, NT_LBO192), dNXlo = desc(NXlo, NT_LBO192);
    float* stg = stage_all + grp * TILE_M * NT_ROW;
    float* myrow = stg + t * NT_ROW;
    float* accS = acc_all + grp * 8;
    int* sg = reinterpret_cast<int*>(accS + 4);
    uint64_t* sbar = bars + grp * 2;
    uint64_t* mbar = bars + grp * 2 + 1;
    const uint32_t bar_id = 1 + grp;
    uint32_t sph = 0, mph = 0;             // mbarrier phase parities
    int cur_graph = -1;
    int it = 0;

    const int64_t num_tiles = (a.N + TILE_M - 1) / TILE_M;
    const int64_t tstride = (int64_t)gridDim.x * NT_GROUPS;
    // TMA: one 256-byte row per node of tile `tl` into the padded staging buffer (no-op past the last tile)
    auto stage_tile = [&](const float* src, int64_t tl) {
        if (tl < num_tiles) {
            const int64_t m0 = tl * TILE_M;
            const int nv = (int)min((int64_t)TILE_M, a.N - m0);
            if (t == 0) mbar_expect_tx(sbar, (uint32_t)nv * (H * 4));
            if (t < nv) bulk_g2s(myrow, src + (size_t)(m0 + t) * H, H * 4, sbar);
        }
    };
    stage_tile(a.h, (int64_t)blockIdx.x * NT_GROUPS + grp);      // first tile; later tiles are prefetched below
    for (int64_t tile = (int64_t)blockIdx.x * NT_GROUPS + grp; tile < num_tiles; tile += tstride, ++it) {
        int* sgp = sg + 2 * (it & 1);                // first/last graph id of the tile, double-buffered by tile parity
        const int64_t n0 = tile * TILE_M;
        const int nvalid = (int)min((int64_t)TILE_M, a.N - n0);
        const bool valid = t < nvalid;
        const size_t node = (size_t)(n0 + (valid ? t : 0));
        auto stage_rows = [&](const float* src) { stage_tile(src, tile); };
        auto staged = [&]() {
            mbar_wait(sbar, sph);
            sph ^= 1;
            __syncwarp();
        };
        auto mma_done = [&]() {
            mbar_wait(mbar, mph);
            mph ^= 1;
            __syncwarp();
            fence_after_sync();
        };
        auto a_ready = [&]() {                       // A operand written by every thread of the group
            wait_st();
            fence_before_sync();
            named_bar(bar_id, NT_GROUP);
        };

        int g = -1;
        float invdeg = 0.f;
        if (valid) {
            g = __ldg(a.batch + node);
            const int deg = __ldg(a.rowptr + node + 1) - __ldg(a.rowptr + node);
            invdeg = 1.0f / (float)max(deg, 1);
        }
        if (t == 0) sgp[0] = g;
        if (t == nvalid - 1) sgp[1] = g;

        // ---- h -> A (fp16 hi/lo) and an fp32 copy in TMEM for the residual -----------------------------------
        staged();
        const float s_h = tc16::encode_row_s(
            [&](int c, float (&v)[16], bool first) {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const float4 x = valid ? *reinterpret_cast<const float4*>(myrow + 16 * c + 4 * j4)
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[4 * j4 + 0] = x.x; v[4 * j4 + 1] = x.y; v[4 * j4 + 2] = x.z; v[4 * j4 + 3] = x.w;
                }
                if (first && !last) {
                    uint32_t raw[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) raw[j] = __float_as_uint(v[j]);
                    tmem_st16(lane_off + tD3 + 16 * c, raw);
                }
            },
            lane_off + tA_hi, lane_off + tA_lo, 1.0f);
        const float inv_h = 1.0f / s_h;
        a_ready();
        if (t == 0) {
            fence_after_sync();
            tc16::issue_f16x3<NT_LBO64>(tD1, tA_hi, tA_lo, dLhi, dLlo, idesc64, false);
            if (!last) tc16::issue_f16x3<NT_LBO64>(tD2, tA_hi, tA_lo, desc(N1hi, NT_LBO64), desc(N1lo, NT_LBO64), idesc64, false);
            mma_commit(mbar);
        }
        __syncwarp();
        if (!last) stage_rows(a.agg_m);              // staging buffer is free: everyone passed the barrier
        else stage_tile(a.h, tile + tstride);        // last layer: h of the next tile right away

        // ---- while MMA 1 runs: per-node coordinate terms --------------------------------------------------
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f), ax = x, tv = x;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f;
        if (valid) {
            x = ldg4(a.x4 + node * 4);
            ax = ldg4(a.agg_x + node * 4);
            if (zero_agg) *reinterpret_cast<float4*>(const_cast<float*>(a.agg_x) + node * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
            tv = ldg4(a.trans_v + node * 4);
            v0 = __ldg(a.vel + node * 3);
            v1 = __ldg(a.vel + node * 3 + 1);
            v2 = __ldg(a.vel + node * 3 + 2);
        }
        const int g_first = sgp[0];                  // written before the barrier above
        const bool single = g_first == sgp[1];
        if (single && g_first != cur_graph) {        // group-uniform
            if (cur_graph >= 0 && t < 4) {
                atomicAdd(a.vsum + (size_t)cur_graph * a.K + t, accS[t]);
                accS[t] = 0.f;
            }
            cur_graph = g_first;
            named_bar(bar_id, NT_GROUP);             // flush complete before new contributions arrive
        }
        mma_done();
        // φ_v(h) (FastEGNN.py:183: the OLD h), then x' = x + agg_x/deg + trans_v + φ_v·v
        float phiv = __ldg(a.lb3);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t d[16];
            tmem_ld16(lane_off + tD1 + 16 * c, d);
            wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j)
                phiv = fmaf(silu(fmaf(__uint_as_float(d[j]), inv_h, lbs[16 * c + j])), lw3s[16 * c + j], phiv);
        }
        float4 xn = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
            xn.x = x.x + ax.x * invdeg + tv.x + phiv * v0;
            xn.y = x.y + ax.y * invdeg + tv.y + phiv * v1;
            xn.z = x.z + ax.z * invdeg + tv.z + phiv * v2;
            *reinterpret_cast<float4*>(a.x4_out + node * 4) = xn;
            if (a.loc_out) {
                a.loc_out[node * 3 + 0] = xn.x;
                a.loc_out[node * 3 + 1] = xn.y;
                a.loc_out[node * 3 + 2] = xn.z;
            }
        }
        if (single) {                                // Σ(x',1) of the tile -> group accumulator
            float s4[4] = {xn.x, xn.y, xn.z, valid ? 1.f : 0.f};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s4[j] += __shfl_xor_sync(FULL, s4[j], o);
            }
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) atomicAdd(accS + j, s4[j]);
            }
        } else if (valid) {
            float* dst = a.vsum + (size_t)g * a.K;
            atomicAdd(dst + 0, xn.x);
            atomicAdd(dst + 1, xn.y);
            atomicAdd(dst + 2, xn.z);
            atomicAdd(dst + 3, 1.0f);
        }
        if (last) {
            fence_before_sync();                     // D1 reads ordered before the next tile's MMA
            continue;
        }

        // ---- node MLP layer 1, chunks 2 and 3: agg_m / deg, agg_v  (D2 accumulates with one row scale) ----------
        float s2 = s_h;                              // scale the D2 row currently carries
        auto l1_chunk = [&](float rs, uint64_t bhi, uint64_t blo, const float* next_src, float* zero_dst) {
            staged();
            if (zero_dst && valid) {                  // the staged row is in shared memory: its source can be cleared
                bulk_s2g(zero_dst + node * H, zero_row, H * 4);
                bulk_commit();
            }
            const float sn = tc16::encode_row_s(
                [&](int c, float (&v)[16], bool) {
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 q = valid ? *reinterpret_cast<const float4*>(myrow + 16 * c + 4 * j4)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
                        v[4 * j4 + 0] = q.x * rs; v[4 * j4 + 1] = q.y * rs; v[4 * j4 + 2] = q.z * rs; v[4 * j4 + 3] = q.w * rs;
                    }
                },
                lane_off + tA_hi, lane_off + tA_lo, s2);
            if (__any_sync(FULL, sn != s2)) {        // cold: bring the partial sums in D2 to the new row scale
                const float r = sn / s2;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t d[16];
                    tmem_ld16(lane_off + tD2 + 16 * c, d);
                    wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) d[j] = __float_as_uint(__uint_as_float(d[j]) * r);
                    tmem_st16(lane_off + tD2 + 16 * c, d);
                }
                s2 = sn;
            }
            a_ready();
            if (t == 0) {
                fence_after_sync();
                tc16::issue_f16x3<NT_LBO64>(tD2, tA_hi, tA_lo, bhi, blo, idesc64, true);
                mma_commit(mbar);
            }
            __syncwarp();
            if (next_src) stage_rows(next_src);
            else stage_tile(a.h, tile + tstride);    // staging buffer is idle for the rest of this tile: prefetch h
            mma_done();
        };
        l1_chunk(invdeg, desc(N1hi + NT_W, NT_LBO64), desc(N1lo + NT_W, NT_LBO64), a.agg_v,
                 zero_agg ? const_cast<float*>(a.agg_m) : nullptr);
        l1_chunk(1.0f, desc(N1hi + 2 * NT_W, NT_LBO64), desc(N1lo + 2 * NT_W, NT_LBO64), nullptr, nullptr);

        // ---- t1 = SiLU(D2/s + attr·N1d + b1) -> A;  D1 = t1·N2ᵀ ----------------------------------------------
        float attrv[DISTEGNN_MAX_NODE_ATTR];
#pragma unroll
        for (int k = 0; k < DISTEGNN_MAX_NODE_ATTR; ++k) attrv[k] = (k < Na && valid) ? __ldg(a.attr + node * Na + k) : 0.f;
        const float inv_s2 = 1.0f / s2;
        const float inv_t = tc16::encode_row(
            [&](int c, float (&v)[16], bool) {
                uint32_t d[16];
                tmem_ld16(lane_off + tD2 + 16 * c, d);
                wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float z = fmaf(__uint_as_float(d[j]), inv_s2, nb1s[16 * c + j]);
#pragma unroll
                    for (int k = 0; k < DISTEGNN_MAX_NODE_ATTR; ++k)
                        if (k < Na) z = fmaf(attrv[k], n1ds[k * H + 16 * c + j], z);
                    v[j] = silu(z);
                }
            },
            lane_off + tA_hi, lane_off + tA_lo);
        a_ready();
        if (t == 0) {
            fence_after_sync();
            tc16::issue_f16x3<NT_LBO64>(tD1, tA_hi, tA_lo, dN2hi, dN2lo, idesc64, false);
            mma_commit(mbar);
        }
        __syncwarp();
        mma_done();

        // ---- h' = h + D1/s + b2 -> HBM and -> A;  [D1|D2|D3] = h'·[W1a';W1b';W1vh']ᵀ -----------------------------
        const float inv_n = tc16::encode_row(
            [&](int c, float (&v)[16], bool first) {
                uint32_t d[16], h0[16];
                tmem_ld16(lane_off + tD1 + 16 * c, d);
                tmem_ld16(lane_off + tD3 + 16 * c, h0);
                wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    v[j] = __uint_as_float(h0[j]) + fmaf(__uint_as_float(d[j]), inv_t, nb2s[16 * c + j]);
                if (first && valid) {
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4)
                        *reinterpret_cast<float4*>(a.h_out + node * H + 16 * c + 4 * j4) =
                            make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
                }
            },
            lane_off + tA_hi, lane_off + tA_lo);
        a_ready();
        if (t == 0) {
            fence_after_sync();
            tc16::issue_f16x3<NT_LBO192>(tD1, tA_hi, tA_lo, dNXhi, dNXlo, idesc192, false);
            mma_commit(mbar);
        }
        __syncwarp();
        mma_done();

        // ---- P = D1/s + b1', Q = D2/s, Hn = D3/s -> HBM ------------------------------------------------------
#pragma unroll 1
        for (int o = 0; o < 3; ++o) {
            float* dst = (o == 0 ? a.P : (o == 1 ? a.Q : a.Hn)) + node * H;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t d[16];
                tmem_ld16(lane_off + tD1 + 64 * o + 16 * c, d);
                wait_ld();
                if (valid) {
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        float4 r;
                        r.x = __uint_as_float(d[4 * j4 + 0]) * inv_n;
                        r.y = __uint_as_float(d[4 * j4 + 1]) * inv_n;
                        r.z = __uint_as_float(d[4 * j4 + 2]) * inv_n;
                        r.w = __uint_as_float(d[4 * j4 + 3]) * inv_n;
                        if (o == 0) {
                            const float4 bb = *reinterpret_cast<const float4*>(nxb1s + 16 * c + 4 * j4);
                            r.x += bb.x; r.y += bb.y; r.z += bb.z; r.w += bb.w;
                        }
                        *reinterpret_cast<float4*>(dst + 16 * c + 4 * j4) = r;
                    }
                }
            }
        }
        fence_before_sync();                         // accumulator reads ordered before the next tile's MMAs
    }
    named_bar(bar_id, NT_GROUP);
    if (cur_graph >= 0 && t < 4) atomicAdd(a.vsum + (size_t)cur_graph * a.K + t, accS[t]);
    if (zero_agg) bulk_wait_all();                   // this thread's zero-row stores have left shared memory

    fence_before_sync();
    __syncthreads();
    if ((tid >> 5) == 0) tmem_dealloc(tbase, 512);
}

// =================================================================================================
// embedding prologue on the tensor cores (FastEGNN.forward, reference models/FastEGNN.py:298-302):
// h0 = embedding_in(node_feat) per node (F <= 16 inputs: plain FMAs), then P/Q/Hn of layer 0 as one N = 192
// kind::f16 tile GEMM; also node_loc -> x4, data_batch -> int32, and Σ(x,1) per graph into vsum.
// =================================================================================================
struct EmbedTcArgs {
    int64_t N;
    int B, F, K;
    const float* feat; const float* loc; const int64_t* batch64;
    const float* wt; const float* bias;                                  // [F][64], [64]
    const float* nw1a; const float* nxb1; const float* nw1b; const float* nw1h;
    float* h; float* x4; int32_t* batch32; float* P; float* Q; float* Hn; float* vsum;
    int32_t* n_invalid;                                                   // device counter of bad data_batch entries (or null)
};
constexpr int ET_SMEM_BYTES = 2 * 3 * NT_W * 2 + (DISTEGNN_MAX_NODE_FEAT + 2) * H * 4 + NT_GROUPS * 8 * 4 + 64;

__global__ void __launch_bounds__(NT_THREADS, 1) embed_tc_kernel(const EmbedTcArgs a) {
    using namespace umma;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __half* NXhi = reinterpret_cast<__half*>(smem_raw);
    __half* NXlo = NXhi + 3 * NT_W;
    float* wts = reinterpret_cast<float*>(NXlo + 3 * NT_W);     // [F][64]
    float* bs = wts + DISTEGNN_MAX_NODE_FEAT * H;
    float* nxb1s = bs + H;
    float* acc_all = nxb1s + H;
    uint64_t* bars = reinterpret_cast<uint64_t*>(acc_all + NT_GROUPS * 8);
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + NT_GROUPS);
    const int tid = threadIdx.x;
    const int grp = tid >> 7, t = tid & 127, lane = tid & 31, wq = (tid >> 5) & 3;
    const int F = a.F;

    tc16::stage_weight<NT_THREADS>(NXhi, NXlo, a.nw1a, 0, 192, tid);
    tc16::stage_weight<NT_THREADS>(NXhi, NXlo, a.nw1b, 64, 192, tid);
    tc16::stage_weight<NT_THREADS>(NXhi, NXlo, a.nw1h, 128, 192, tid);
    for (int i = tid; i < F * H; i += NT_THREADS) wts[i] = a.wt[i];
    if (tid < H) {
        bs[tid] = a.bias[tid];
        nxb1s[tid] = a.nxb1[tid];
    }
    if (tid < NT_GROUPS * 8) acc_all[tid] = 0.f;
    if (tid == 0) {
        for (int i = 0; i < NT_GROUPS; ++i) mbar_init(&bars[i], 1);
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
    const uint32_t col0 = tbase + (uint32_t)grp * 256u;
    const uint32_t tA_hi = col0, tA_lo = col0 + 32, tD1 = col0 + 64;
    const uint32_t idesc192 = make_idesc_f16(128, 192, 0, 0);
    const uint64_t dNXhi = make_b_desc(smem_u32(NXhi), NT_LBO192, 128), dNXlo = make_b_desc(smem_u32(NXlo), NT_LBO192, 128);
    float* accS = acc_all + grp * 8;
    int* sg = reinterpret_cast<int*>(accS + 4);
    uint64_t* mbar = bars + grp;
    const uint32_t bar_id = 1 + grp;
    uint32_t mph = 0;
    int cur_graph = -1, it = 0;

    const int64_t num_tiles = (a.N + TILE_M - 1) / TILE_M;
    for (int64_t tile = (int64_t)blockIdx.x * NT_GROUPS + grp; tile < num_tiles; tile += (int64_t)gridDim.x * NT_GROUPS, ++it) {
        int* sgp = sg + 2 * (it & 1);
        const int64_t n0 = tile * TILE_M;
        const int nvalid = (int)min((int64_t)TILE_M, a.N - n0);
        const bool valid = t < nvalid;
        const size_t node = (size_t)(n0 + (valid ? t : 0));
        int g = -1;
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
        float f[DISTEGNN_MAX_NODE_FEAT];
#pragma unroll
        for (int k = 0; k < DISTEGNN_MAX_NODE_FEAT; ++k) f[k] = (valid && k < F) ? __ldg(a.feat + node * F + k) : 0.f;
        if (valid) {
            // precondition of every per-graph reduction downstream: ids sorted and inside [0,B) (PyG batches are; the
            // reference takes B from data_batch[-1]+1, FastEGNN.py:298).  Violations are counted for the host and clamped.
            const int64_t gi = a.batch64[node];
            const bool bad = gi < 0 || gi >= a.B || (node > 0 && a.batch64[node - 1] > gi);
            if (bad && a.n_invalid) atomicAdd(a.n_invalid, 1);
            g = (int)(gi < 0 ? 0 : (gi >= a.B ? a.B - 1 : gi));
            a.batch32[node] = g;
            const float* p = a.loc + node * 3;
            xv = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
            *reinterpret_cast<float4*>(a.x4 + node * 4) = xv;
        }
        if (t == 0) sgp[0] = g;
        if (t == nvalid - 1) sgp[1] = g;

        // h0 row -> HBM and -> A
        const float inv_n = tc16::encode_row(
            [&](int c, float (&v)[16], bool first) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float z = bs[16 * c + j];
#pragma unroll
                    for (int k = 0; k < DISTEGNN_MAX_NODE_FEAT; ++k)
                        if (k < F) z = fmaf(f[k], wts[k * H + 16 * c + j], z);
                    v[j] = valid ? z : 0.f;
                }
                if (first && valid) {
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4)
                        *reinterpret_cast<float4*>(a.h + node * H + 16 * c + 4 * j4) =
                            make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
                }
            },
            lane_off + tA_hi, lane_off + tA_lo);
        wait_st();
        fence_before_sync();
        named_bar(bar_id, NT_GROUP);
        if (t == 0) {
            fence_after_sync();
            tc16::issue_f16x3<NT_LBO192>(tD1, tA_hi, tA_lo, dNXhi, dNXlo, idesc192, false);
            mma_commit(mbar);
        }
        __syncwarp();
        // Σ(x,1) per graph while the MMA runs
        const int g_first = sgp[0];
        const bool single = g_first == sgp[1];
        if (single && g_first != cur_graph) {
            if (cur_graph >= 0 && t < 4) {
                atomicAdd(a.vsum + (size_t)cur_graph * a.K + t, accS[t]);
                accS[t] = 0.f;
            }
            cur_graph = g_first;
            named_bar(bar_id, NT_GROUP);
        }
        if (single) {
            float s4[4] = {xv.x, xv.y, xv.z, valid ? 1.f : 0.f};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s4[j] += __shfl_xor_sync(FULL, s4[j], o);
            }
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) atomicAdd(accS + j, s4[j]);
            }
        } else if (valid) {
            float* dst = a.vsum + (size_t)g * a.K;
            atomicAdd(dst + 0, xv.x);
            atomicAdd(dst + 1, xv.y);
            atomicAdd(dst + 2, xv.z);
            atomicAdd(dst + 3, 1.0f);
        }
        mbar_wait(mbar, mph);
        mph ^= 1;
        __syncwarp();
        fence_after_sync();
#pragma unroll 1
        for (int o = 0; o < 3; ++o) {
            float* dst = (o == 0 ? a.P : (o == 1 ? a.Q : a.Hn)) + node * H;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t d[16];
                tmem_ld16(lane_off + tD1 + 64 * o + 16 * c, d);
                wait_ld();
                if (valid) {
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        float4 r;
                        r.x = __uint_as_float(d[4 * j4 + 0]) * inv_n;
                        r.y = __uint_as_float(d[4 * j4 + 1]) * inv_n;
                        r.z = __uint_as_float(d[4 * j4 + 2]) * inv_n;
                        r.w = __uint_as_float(d[4 * j4 + 3]) * inv_n;
                        if (o == 0) {
                            const float4 bb = *reinterpret_cast<const float4*>(nxb1s + 16 * c + 4 * j4);
                            r.x += bb.x; r.y += bb.y; r.z += bb.z; r.w += bb.w;
                        }
                        *reinterpret_cast<float4*>(dst + 16 * c + 4 * j4) = r;
                    }
                }
            }
        }
        fence_before_sync();
    }
    named_bar(bar_id, NT_GROUP);
    if (cur_graph >= 0 && t < 4) atomicAdd(a.vsum + (size_t)cur_graph * a.K + t, accS[t]);
    fence_before_sync();
    __syncthreads();
    if ((tid >> 5) == 0) tmem_dealloc(tbase, 512);
}

}  // namespace degnn

extern "C" int distegnn_node_layer_fwd(int64_t n_nodes, int n_graphs, int A, int C, int Na, unsigned flags,
                                       const int32_t* rowptr, const int32_t* batch32, const float* h, const float* x4,
                                       const float* node_vel, const float* node_attr, const float* agg_m,
                                       const float* agg_x, const float* agg_v, const float* trans_v,
                                       const float* layer_params, const float* next_layer_params, float* h_out,
                                       float* x4_out, float* P, float* Q, float* Hn, float* node_loc_out, float* vsum,
                                       void* stream) {
    using namespace degnn;
    if (int rc = check_dims(A, C, Na)) return rc;
    if (n_nodes == 0) return DISTEGNN_OK;
    const bool last = flags & DISTEGNN_FLAG_LAST;
    DEGNN_CHECK_ARG(n_nodes > 0 && n_graphs > 0, "bad size");
    DEGNN_CHECK_ARG(rowptr && batch32 && h && x4 && node_vel && agg_x && trans_v && layer_params && x4_out && vsum,
                    "null pointer");
    DEGNN_CHECK_ARG(Na == 0 || last || node_attr, "null node_attr with node_attr_nf > 0");
    DEGNN_CHECK_ARG(last || (agg_m && agg_v && next_layer_params && h_out && P && Q && Hn),
                    "null pointer (non-last layer)");
    Layout L = make_layout(A, C, Na);
    NodeTcArgs a;
    a.N = n_nodes; a.B = n_graphs; a.Na = Na; a.K = 4 + 3 * C + H * C; a.flags = flags;
    a.rowptr = rowptr; a.batch = batch32; a.h = h; a.x4 = x4; a.vel = node_vel; a.attr = node_attr;
    a.agg_m = agg_m; a.agg_x = agg_x; a.agg_v = agg_v; a.trans_v = trans_v;
    a.lw = layer_params + L.off[DISTEGNN_P_L_W];
    a.lb = layer_params + L.off[DISTEGNN_P_L_B];
    a.lw3 = layer_params + L.off[DISTEGNN_P_L_W3];
    a.lb3 = layer_params + L.off[DISTEGNN_P_L_B3];
    a.n1 = layer_params + L.off[DISTEGNN_P_N_W1];
    a.nb1 = layer_params + L.off[DISTEGNN_P_N_B1];
    a.n2 = layer_params + L.off[DISTEGNN_P_N_W2];
    a.nb2 = layer_params + L.off[DISTEGNN_P_N_B2];
    const float* nx = next_layer_params ? next_layer_params : layer_params;
    a.nw1a = nx + L.off[DISTEGNN_P_E_W1A];
    a.nxb1 = nx + L.off[DISTEGNN_P_E_B1];
    a.nw1b = nx + L.off[DISTEGNN_P_E_W1B];
    a.nw1h = nx + L.off[DISTEGNN_P_V_W1H];
    a.h_out = h_out; a.x4_out = x4_out; a.P = P; a.Q = Q; a.Hn = Hn; a.loc_out = node_loc_out; a.vsum = vsum;
    ensure_dynamic_smem((const void*)node_layer_tc_kernel, (int)NT_SMEM_BYTES);
    const int64_t tiles = (n_nodes + TILE_M - 1) / TILE_M;
    int64_t grid = (tiles + NT_GROUPS - 1) / NT_GROUPS;
    if (grid > sm_count()) grid = sm_count();
    node_layer_tc_kernel<<<(unsigned)grid, NT_THREADS, NT_SMEM_BYTES, (cudaStream_t)stream>>>(a);
    DEGNN_CHECK_LAUNCH();
    return DISTEGNN_OK;
}

extern "C" int distegnn_embed_fwd(int64_t n_nodes, int n_graphs, int F, int A, int C, int Na, const float* node_feat,
                                  const float* node_loc, const int64_t* data_batch, const float* emb_wt,
                                  const float* emb_b, const float* layer0_params, float* h, float* x4,
                                  int32_t* batch32, float* P, float* Q, float* Hn, float* vsum, int32_t* n_invalid,
                                  void* stream) {
    using namespace degnn;
    if (int rc = check_dims(A, C, Na)) return rc;
    if (n_nodes == 0) return DISTEGNN_OK;
    DEGNN_CHECK_ARG(n_nodes > 0 && n_graphs > 0, "bad size");
    DEGNN_CHECK_ARG(F >= 1 && F <= DISTEGNN_MAX_NODE_FEAT, "node_feat_nf out of range");
    DEGNN_CHECK_ARG(node_feat && node_loc && data_batch && emb_wt && emb_b && layer0_params && h && x4 && batch32 &&
                        P && Q && Hn && vsum, "null pointer");
    Layout L = make_layout(A, C, Na);
    EmbedTcArgs a;
    a.N = n_nodes; a.B = n_graphs; a.F = F; a.K = 4 + 3 * C + H * C;
    a.feat = node_feat; a.loc = node_loc; a.batch64 = data_batch; a.wt = emb_wt; a.bias = emb_b;
    a.nw1a = layer0_params + L.off[DISTEGNN_P_E_W1A];
    a.nxb1 = layer0_params + L.off[DISTEGNN_P_E_B1];
    a.nw1b = layer0_params + L.off[DISTEGNN_P_E_W1B];
    a.nw1h = layer0_params + L.off[DISTEGNN_P_V_W1H];
    a.h = h; a.x4 = x4; a.batch32 = batch32; a.P = P; a.Q = Q; a.Hn = Hn; a.vsum = vsum;
    a.n_invalid = n_invalid;
    ensure_dynamic_smem((const void*)embed_tc_kernel, (int)ET_SMEM_BYTES);
    const int64_t tiles = (n_nodes + TILE_M - 1) / TILE_M;
    int64_t grid = (tiles + NT_GROUPS - 1) / NT_GROUPS;
    if (grid > sm_count()) grid = sm_count();
    embed_tc_kernel<<<(unsigned)grid, NT_THREADS, ET_SMEM_BYTES, (cudaStream_t)stream>>>(a);
    DEGNN_CHECK_LAUNCH();
    return DISTEGNN_OK;
}
