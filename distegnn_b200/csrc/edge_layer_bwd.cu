// Backward of the real<->real edge stage — tile GEMMs as 3xTF32 warp MMAs (mma_tf32.cuh).
// Differentiates what distegnn_edge_layer_fwd computes (reference models/FastEGNN.py:237-246 coord2radial,
// :144-150 edge_model, :169-177 edge part of coord_model_vel, :206 edge part of node_model, scatter_add_ :322-337;
// in the reference this is autograd through ~20 [E,64] tensors):
//     z1 = P[i] + Q[j] + w_r·r + W_e·a,  a1 = SiLU(z1),  z2 = W2·a1 + b2,  m = SiLU(z2),
//     zc = Wc·m + bc,  φ = w3·SiLU(zc),   agg_m[i] += m,   agg_x[i] += Δ·φ        (i = row, j = col)
// Nothing of size [E,·] is kept from the forward pass: each 128-edge tile is recomputed (2 tile GEMMs), then
//     gφ = g_aggx[i]·Δ,  g_zc = gφ·w3 ⊙ SiLU'(zc),  g_m = g_aggm[i] + Wcᵀ·g_zc,  g_z2 = g_m ⊙ SiLU'(z2),
//     g_z1 = (W2ᵀ·g_z2) ⊙ SiLU'(z1)                                               (2 tile GEMMs)
//     g_P[i] += g_z1,  g_Q[j] += g_z1,  g_r = w_r·g_z1,  gΔ_raw = g_aggx[i]·φ/norm + 2·g_r·Δ_raw,
//     g_x[i] += gΔ_raw,  g_x[j] −= gΔ_raw                                          (norm detached, :243)
// and the weight gradients g_Wc += g_zcᵀ·m, g_W2 += g_z2ᵀ·a1 (2 more tile GEMMs, accumulated in registers over all
// tiles of the CTA), g_b2, g_bc, g_w3, g_w_r, g_W_e (accumulated in shared memory), all added to `g_params`, a
// buffer with the layout of the parameter block (distegnn_param_layout).  The weights are read from global memory (L1);
// the transposed matrices of the data-gradient GEMMs come in `wT`, built by the caller.
#include "bwd_common.cuh"
#include "common.cuh"
#include "mma_tf32.cuh"

namespace degnn {

struct EdgeBwdArgs {
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
    const float* w2;
    const float* b2;
    const float* wc;
    const float* bc;
    const float* w3;
    const float* wT;       // [2][64][64]: W2ᵀ, Wcᵀ (wT[m][n*64+k] = W_m[k*64+n])
    const float* g_aggm;   // [N,64] gradient w.r.t. the SUM agg_m (null with FLAG_LAST)
    const float* g_aggx;   // [N,4]  gradient w.r.t. the SUM agg_x
    float* g_P;            // [N,64] +=
    float* g_Q;            // [N,64] +=
    float* g_x;            // [N,4]  +=
    float* g_w1r; float* g_w1e; float* g_w2; float* g_b2; float* g_wc; float* g_bc; float* g_w3;   // += (param layout)
};

constexpr int EB_SMEM_FLOATS = 4 * TILE_M * LDA          // Z1, Z2, W (A operand), G (gradient operand)
                               + 4 * H + DISTEGNN_MAX_EDGE_ATTR * H          // b2, bc, w3, w1r, w1e
                               + 4 * H + DISTEGNN_MAX_EDGE_ATTR * H          // gradient accumulators of the same
                               + TILE_M * DISTEGNN_MAX_EDGE_ATTR             // edge attrs of the tile
                               + TILE_M * 4              // Δ_raw (xyz) + radial
                               + 6 * TILE_M              // 1/norm, gφ, φ, g_r, row, col
                               + 4;                      // run-start masks (one word per 32 edges)
constexpr size_t EB_SMEM_BYTES = EB_SMEM_FLOATS * sizeof(float);

__global__ void __launch_bounds__(NTHREADS, 1) edge_layer_bwd_kernel(const EdgeBwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* Z1 = smem;
    float* Z2 = Z1 + TILE_M * LDA;
    float* Wt = Z2 + TILE_M * LDA;          // current A operand: a1, then m, then a1 again
    float* Gt = Wt + TILE_M * LDA;          // current gradient operand: g_zc, then g_z2
    float* b2s = Gt + TILE_M * LDA;
    float* bcs = b2s + H;
    float* w3s = bcs + H;
    float* w1rs = w3s + H;
    float* w1es = w1rs + H;
    float* gb2 = w1es + DISTEGNN_MAX_EDGE_ATTR * H;
    float* gbc = gb2 + H;
    float* gw3 = gbc + H;
    float* gw1r = gw3 + H;
    float* gw1e = gw1r + H;
    float* eas = gw1e + DISTEGNN_MAX_EDGE_ATTR * H;
    float* dxs = eas + TILE_M * DISTEGNN_MAX_EDGE_ATTR;      // Δ_raw xyz, radial
    float* invn = dxs + TILE_M * 4;
    float* gphis = invn + TILE_M;
    float* phis = gphis + TILE_M;
    float* grs = phis + TILE_M;
    int* srow = reinterpret_cast<int*>(grs + TILE_M);
    int* scol = srow + TILE_M;
    uint32_t* rmask = reinterpret_cast<uint32_t*>(scol + TILE_M);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int R0 = 16 * warp + g, R1 = R0 + 8;          // the two tile rows of this thread's accumulator fragments
    const int A = a.A;
    const bool normalize = a.flags & DISTEGNN_FLAG_NORMALIZE;
    const bool need_m = !(a.flags & DISTEGNN_FLAG_LAST) && a.g_aggm != nullptr;
    const float* W2T = a.wT;
    const float* WcT = a.wT + H * H;

    if (tid < H) {
        b2s[tid] = a.b2[tid];
        bcs[tid] = a.bc[tid];
        w3s[tid] = a.w3[tid];
        w1rs[tid] = a.w1r[tid];
    }
    for (int i = tid; i < A * H; i += NTHREADS) w1es[i] = a.w1e[i];
    for (int i = tid; i < 4 * H + DISTEGNN_MAX_EDGE_ATTR * H; i += NTHREADS) gb2[i] = 0.f;
    float gW2[4][4], gWc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) gW2[i][j] = gWc[i][j] = 0.f;
    __syncthreads();

    const int64_t num_tiles = (a.E + TILE_M - 1) / TILE_M;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t e0 = tile * TILE_M;
        // ---- 1. edge metadata -----------------------------------------------------------------------------------
        if (tid < TILE_M) {
            const int64_t e = e0 + tid;
            const bool ok = e < a.E;
            srow[tid] = ok ? __ldg(a.row + e) : -1;
            scol[tid] = ok ? __ldg(a.col + e) : 0;
        }
        for (int i = tid; i < TILE_M * A; i += NTHREADS) {
            const int64_t gi = e0 * A + i;
            eas[i] = (gi < a.E * A) ? __ldg(a.ea + gi) : 0.f;
        }
        __syncthreads();
        if (tid < TILE_M) {      // bit i of rmask[q] = edge 32q+i starts a new run of equal destination rows
            const int prev = tid > 0 ? srow[tid - 1] : -2;
            const uint32_t starts = __ballot_sync(FULL, prev != srow[tid]);
            if (lane == 0) rmask[warp] = starts;
        }

        // ---- 2. gather + first layer (half-warp per edge): Z1 = z1, Wt = a1, geometry, gφ ----------------------------
        {
            const int l = lane & 15;
            const float4 wr4 = *reinterpret_cast<const float4*>(w1rs + 4 * l);
#pragma unroll 2
            for (int it = 0; it < 8; ++it) {
                const int el = 16 * warp + 2 * it + (lane >> 4);
                const int r = srow[el];
                float4 pre = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r >= 0) {
                    const int c = scol[el];
                    const float4 p = ldg4(a.P + (size_t)r * H + 4 * l);
                    const float4 q = ldg4(a.Q + (size_t)c * H + 4 * l);
                    const float4 xi = ldg4(a.x4 + (size_t)r * 4);
                    const float4 xj = ldg4(a.x4 + (size_t)c * 4);
                    const float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
                    const float radial = dx * dx + dy * dy + dz * dz;
                    pre = fma4(radial, wr4, add4(p, q));
                    for (int k = 0; k < A; ++k)
                        pre = fma4(eas[el * A + k], *reinterpret_cast<const float4*>(w1es + k * H + 4 * l), pre);
                    if (l == 0) {
                        const float inv = normalize ? 1.0f / (sqrtf(radial) + 1e-8f) : 1.0f;
                        const float4 gx = ldg4(a.g_aggx + (size_t)r * 4);
                        *reinterpret_cast<float4*>(dxs + 4 * el) = make_float4(dx, dy, dz, radial);
                        invn[el] = inv;
                        gphis[el] = (gx.x * dx + gx.y * dy + gx.z * dz) * inv;
                    }
                } else if (l == 0) {
                    *reinterpret_cast<float4*>(dxs + 4 * el) = make_float4(0.f, 0.f, 0.f, 0.f);
                    invn[el] = 0.f;
                    gphis[el] = 0.f;
                }
                *reinterpret_cast<float4*>(Z1 + el * LDA + 4 * l) = pre;
                *reinterpret_cast<float4*>(Wt + el * LDA + 4 * l) = silu4(pre);
            }
        }
        __syncthreads();

        // ---- 3. z2 = a1·W2 + b2 -> Z2;  Wt = m = SiLU(z2) --------------------------------------------------------
        float acc[8][4];
        mma3::zero(acc);
        mma3::gemm_rows<true>(acc, Wt, a.w2, H, warp, lane);
        __syncthreads();                       // everyone finished reading a1
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = 8 * j + 2 * t;
            const float2 bb = *reinterpret_cast<const float2*>(b2s + c);
            const float2 z0 = make_float2(acc[j][0] + bb.x, acc[j][1] + bb.y), z1 = make_float2(acc[j][2] + bb.x, acc[j][3] + bb.y);
            *reinterpret_cast<float2*>(Z2 + R0 * LDA + c) = z0;
            *reinterpret_cast<float2*>(Z2 + R1 * LDA + c) = z1;
            *reinterpret_cast<float2*>(Wt + R0 * LDA + c) = make_float2(silu(z0.x), silu(z0.y));
            *reinterpret_cast<float2*>(Wt + R1 * LDA + c) = make_float2(silu(z1.x), silu(z1.y));
        }
        __syncthreads();

        // ---- 4. zc = m·Wc + bc;  φ;  g_zc = gφ·w3 ⊙ SiLU'(zc) -> Gt;  g_w3, g_bc ---------------------------------------
        mma3::zero(acc);
        mma3::gemm_rows<true>(acc, Wt, a.wc, H, warp, lane);
        {
            const float gp0 = gphis[R0], gp1 = gphis[R1];
            float ph0 = 0.f, ph1 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = 8 * j + 2 * t;
                const float2 bb = *reinterpret_cast<const float2*>(bcs + c);
                const float2 ww = *reinterpret_cast<const float2*>(w3s + c);
                const float zc[4] = {acc[j][0] + bb.x, acc[j][1] + bb.y, acc[j][2] + bb.x, acc[j][3] + bb.y};
                const float w3a[4] = {ww.x, ww.y, ww.x, ww.y};
                const float gpa[4] = {gp0, gp0, gp1, gp1};
                float gz[4], ac[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float s = sigmoid_f(zc[i]);
                    ac[i] = zc[i] * s;
                    gz[i] = gpa[i] * w3a[i] * (s * fmaf(zc[i], 1.0f - s, 1.0f));
                }
                ph0 = fmaf(ac[0], ww.x, fmaf(ac[1], ww.y, ph0));
                ph1 = fmaf(ac[2], ww.x, fmaf(ac[3], ww.y, ph1));
                atomicAdd(gw3 + c, fmaf(gp0, ac[0], gp1 * ac[2]));
                atomicAdd(gw3 + c + 1, fmaf(gp0, ac[1], gp1 * ac[3]));
                atomicAdd(gbc + c, gz[0] + gz[2]);
                atomicAdd(gbc + c + 1, gz[1] + gz[3]);
                *reinterpret_cast<float2*>(Gt + R0 * LDA + c) = make_float2(gz[0], gz[1]);
                *reinterpret_cast<float2*>(Gt + R1 * LDA + c) = make_float2(gz[2], gz[3]);
            }
            ph0 += __shfl_xor_sync(FULL, ph0, 1);
            ph0 += __shfl_xor_sync(FULL, ph0, 2);
            ph1 += __shfl_xor_sync(FULL, ph1, 1);
            ph1 += __shfl_xor_sync(FULL, ph1, 2);
            if (t == 0) {
                phis[R0] = ph0;
                phis[R1] = ph1;
            }
        }
        __syncthreads();

        // ---- 5. g_Wc += g_zcᵀ·m;  g_m = g_zc·Wc + g_aggm[row];  g_z2 = g_m ⊙ SiLU'(z2) ------------------------------
        mma3::wgrad(gWc, Gt, Wt, warp, lane);
        mma3::zero(acc);
        mma3::gemm_rows<true>(acc, Gt, WcT, H, warp, lane);
        __syncthreads();                       // Gt and Wt fully read
        {
            const int r0 = srow[R0], r1 = srow[R1];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = 8 * j + 2 * t;
                float2 gm0 = make_float2(acc[j][0], acc[j][1]), gm1 = make_float2(acc[j][2], acc[j][3]);
                if (need_m) {
                    if (r0 >= 0) {
                        const float2 u = __ldg(reinterpret_cast<const float2*>(a.g_aggm + (size_t)r0 * H + c));
                        gm0.x += u.x; gm0.y += u.y;
                    }
                    if (r1 >= 0) {
                        const float2 u = __ldg(reinterpret_cast<const float2*>(a.g_aggm + (size_t)r1 * H + c));
                        gm1.x += u.x; gm1.y += u.y;
                    }
                }
                const float2 z0 = *reinterpret_cast<const float2*>(Z2 + R0 * LDA + c);
                const float2 z1 = *reinterpret_cast<const float2*>(Z2 + R1 * LDA + c);
                const float2 g0 = make_float2(gm0.x * dsilu(z0.x), gm0.y * dsilu(z0.y));
                const float2 g1 = make_float2(gm1.x * dsilu(z1.x), gm1.y * dsilu(z1.y));
                atomicAdd(gb2 + c, g0.x + g1.x);
                atomicAdd(gb2 + c + 1, g0.y + g1.y);
                *reinterpret_cast<float2*>(Gt + R0 * LDA + c) = g0;
                *reinterpret_cast<float2*>(Gt + R1 * LDA + c) = g1;
            }
        }
        silu_tile(Wt, Z1, tid);                // Wt = a1 again
        __syncthreads();

        // ---- 6. g_W2 += g_z2ᵀ·a1;  g_z1 = (g_z2·W2) ⊙ SiLU'(z1) -> Z2 tile (dead);  g_r ----------------------------------
        mma3::wgrad(gW2, Gt, Wt, warp, lane);
        mma3::zero(acc);
        mma3::gemm_rows<true>(acc, Gt, W2T, H, warp, lane);
        {
            float gr0 = 0.f, gr1 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = 8 * j + 2 * t;
                const float2 z0 = *reinterpret_cast<const float2*>(Z1 + R0 * LDA + c);
                const float2 z1 = *reinterpret_cast<const float2*>(Z1 + R1 * LDA + c);
                const float2 wr = *reinterpret_cast<const float2*>(w1rs + c);
                const float2 g0 = make_float2(acc[j][0] * dsilu(z0.x), acc[j][1] * dsilu(z0.y));
                const float2 g1 = make_float2(acc[j][2] * dsilu(z1.x), acc[j][3] * dsilu(z1.y));
                gr0 = fmaf(g0.x, wr.x, fmaf(g0.y, wr.y, gr0));
                gr1 = fmaf(g1.x, wr.x, fmaf(g1.y, wr.y, gr1));
                *reinterpret_cast<float2*>(Z2 + R0 * LDA + c) = g0;
                *reinterpret_cast<float2*>(Z2 + R1 * LDA + c) = g1;
            }
            gr0 += __shfl_xor_sync(FULL, gr0, 1);
            gr0 += __shfl_xor_sync(FULL, gr0, 2);
            gr1 += __shfl_xor_sync(FULL, gr1, 1);
            gr1 += __shfl_xor_sync(FULL, gr1, 2);
            if (t == 0) {
                grs[R0] = gr0;
                grs[R1] = gr1;
            }
        }
        __syncthreads();

        // ---- 7. scatter g_z1 (Z2 tile): g_P by runs of equal row, g_Q per edge, g_w_r / g_W_e column sums; geometry ------
        {   // g_P: warp <-> 16 edges, lane <-> column pair, one RED.v2 per run
            const float* colp = Z2 + (16 * warp) * LDA + 2 * lane;
            uint32_t M = ((rmask[warp >> 1] >> (16 * (warp & 1))) & 0xffffu) | 1u;
            while (M) {
                const int s0 = __ffs((int)M) - 1;
                M &= M - 1;
                const int s1 = M ? __ffs((int)M) - 1 : 16;
                float2 s = make_float2(0.f, 0.f);
                for (int e = s0; e < s1; ++e) {
                    const float2 v = *reinterpret_cast<const float2*>(colp + e * LDA);
                    s.x += v.x; s.y += v.y;
                }
                const int r = srow[16 * warp + s0];
                if (r >= 0) {
                    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(a.g_P + (size_t)r * H + 2 * lane), "f"(s.x), "f"(s.y)
                                 : "memory");
                }
            }
        }
        {   // g_Q: half-warp per edge, RED.v4
            const int l = lane & 15;
#pragma unroll 2
            for (int it = 0; it < 8; ++it) {
                const int el = 16 * warp + 2 * it + (lane >> 4);
                if (srow[el] >= 0)
                    red_add_v4(a.g_Q + (size_t)scol[el] * H + 4 * l, *reinterpret_cast<const float4*>(Z2 + el * LDA + 4 * l));
            }
        }
        {   // g_w_r[n] += Σ_e g_z1[e][n]·radial_e,  g_W_e[k][n] += Σ_e g_z1[e][n]·a_ek : thread <-> (column, quarter of the tile)
            const int c = tid & 63, q = tid >> 6;
            float sr = 0.f, se[DISTEGNN_MAX_EDGE_ATTR];
#pragma unroll
            for (int k = 0; k < DISTEGNN_MAX_EDGE_ATTR; ++k) se[k] = 0.f;
            for (int e = 32 * q; e < 32 * q + 32; ++e) {
                const float v = Z2[e * LDA + c];
                sr = fmaf(v, dxs[4 * e + 3], sr);
#pragma unroll
                for (int k = 0; k < DISTEGNN_MAX_EDGE_ATTR; ++k)
                    if (k < A) se[k] = fmaf(v, eas[e * A + k], se[k]);
            }
            atomicAdd(gw1r + c, sr);
#pragma unroll
            for (int k = 0; k < DISTEGNN_MAX_EDGE_ATTR; ++k)
                if (k < A) atomicAdd(gw1e + k * H + c, se[k]);
        }
        if (tid < TILE_M) {      // geometry: gΔ_raw = g_aggx[i]·φ/norm + 2·g_r·Δ_raw
            const int r = srow[tid];
            if (r >= 0) {
                const int c = scol[tid];
                const float4 d = *reinterpret_cast<const float4*>(dxs + 4 * tid);
                const float4 gx = ldg4(a.g_aggx + (size_t)r * 4);
                const float s = phis[tid] * invn[tid], t2 = 2.0f * grs[tid];
                const float gdx = fmaf(gx.x, s, t2 * d.x), gdy = fmaf(gx.y, s, t2 * d.y), gdz = fmaf(gx.z, s, t2 * d.z);
                atomicAdd(a.g_x + (size_t)r * 4 + 0, gdx);
                atomicAdd(a.g_x + (size_t)r * 4 + 1, gdy);
                atomicAdd(a.g_x + (size_t)r * 4 + 2, gdz);
                atomicAdd(a.g_x + (size_t)c * 4 + 0, -gdx);
                atomicAdd(a.g_x + (size_t)c * 4 + 1, -gdy);
                atomicAdd(a.g_x + (size_t)c * 4 + 2, -gdz);
            }
        }
        __syncthreads();                       // tile buffers are rewritten by the next iteration
    }

    // ---- flush the CTA's parameter gradients ------------------------------------------------------------------------
    mma3::wgrad_flush(a.g_w2, gW2, warp, lane);
    mma3::wgrad_flush(a.g_wc, gWc, warp, lane);
    if (tid < H) {
        atomicAdd(a.g_b2 + tid, gb2[tid]);
        atomicAdd(a.g_bc + tid, gbc[tid]);
        atomicAdd(a.g_w3 + tid, gw3[tid]);
        atomicAdd(a.g_w1r + tid, gw1r[tid]);
    }
    for (int i = tid; i < A * H; i += NTHREADS) atomicAdd(a.g_w1e + i, gw1e[i]);
}

}  // namespace degnn

extern "C" int distegnn_edge_layer_bwd(int64_t n_nodes, int64_t n_edges, int A, int C, int Na, unsigned flags,
                                       const int32_t* row, const int32_t* col, const float* edge_attr_sorted,
                                       const float* x4, const float* P, const float* Q, const float* layer_params,
                                       const float* wT, const float* g_agg_m, const float* g_agg_x, float* g_P, float* g_Q, float* g_x4,
                                       float* g_layer_params, void* stream) {
    using namespace degnn;
    if (int rc = check_dims(A, C, Na)) return rc;
    if (n_edges == 0) return DISTEGNN_OK;
    DEGNN_CHECK_ARG(n_nodes > 0 && n_edges > 0, "negative size");
    DEGNN_CHECK_ARG(row && col && x4 && P && Q && layer_params && wT && g_agg_x && g_P && g_Q && g_x4 && g_layer_params,
                    "null pointer");
    DEGNN_CHECK_ARG(A == 0 || edge_attr_sorted, "null edge_attr with edge_attr_nf > 0");
    Layout L = make_layout(A, C, Na);
    EdgeBwdArgs a;
    a.N = n_nodes; a.E = n_edges; a.A = A; a.flags = flags;
    a.row = row; a.col = col; a.ea = edge_attr_sorted; a.x4 = x4; a.P = P; a.Q = Q;
    a.w1r = layer_params + L.off[DISTEGNN_P_E_W1R];
    a.w1e = layer_params + L.off[DISTEGNN_P_E_W1E];
    a.w2 = layer_params + L.off[DISTEGNN_P_E_W2];
    a.b2 = layer_params + L.off[DISTEGNN_P_E_B2];
    a.wc = layer_params + L.off[DISTEGNN_P_E_WC];
    a.bc = layer_params + L.off[DISTEGNN_P_E_BC];
    a.w3 = layer_params + L.off[DISTEGNN_P_E_W3];
    a.wT = wT;
    a.g_aggm = g_agg_m; a.g_aggx = g_agg_x; a.g_P = g_P; a.g_Q = g_Q; a.g_x = g_x4;
    a.g_w1r = g_layer_params + L.off[DISTEGNN_P_E_W1R];
    a.g_w1e = g_layer_params + L.off[DISTEGNN_P_E_W1E];
    a.g_w2 = g_layer_params + L.off[DISTEGNN_P_E_W2];
    a.g_b2 = g_layer_params + L.off[DISTEGNN_P_E_B2];
    a.g_wc = g_layer_params + L.off[DISTEGNN_P_E_WC];
    a.g_bc = g_layer_params + L.off[DISTEGNN_P_E_BC];
    a.g_w3 = g_layer_params + L.off[DISTEGNN_P_E_W3];
    ensure_dynamic_smem((const void*)edge_layer_bwd_kernel, (int)EB_SMEM_BYTES);
    const int64_t tiles = (n_edges + TILE_M - 1) / TILE_M;
    int64_t grid = sm_count();
    if (grid > tiles) grid = tiles;
    edge_layer_bwd_kernel<<<(unsigned)grid, NTHREADS, EB_SMEM_BYTES, (cudaStream_t)stream>>>(a);
    DEGNN_CHECK_LAUNCH();
    return DISTEGNN_OK;
}
