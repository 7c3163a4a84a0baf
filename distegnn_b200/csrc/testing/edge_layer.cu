// Real<->real edge stage of one E_GCL_vel layer, fused: gather -> edge MLP -> coord head ->
// per-destination segment sums.  Replaces coord2radial + edge_model + the edge halves of
// coord_model_vel / node_model (reference models/FastEGNN.py:237-246, 144-150, 169-177, 206) and the
// scatter_add_ of unsorted_segment_sum/mean (:322-337, twins models/basic.py:22-66).
//
// Work decomposition: persistent CTAs, each looping over tiles of 128 consecutive CSR edges
// (sorted by destination row).  Per tile:
//   1. coalesced load of row/col/edge_attr for the 128 edges;
//   2. half-warp per edge: 16-byte loads of P[row] and Q[col] (256 B rows), Δx from x4, first-layer
//      pre-activation P_i + Q_j + w_r·r + W_e·a (the 131-wide concat of the reference is never
//      formed: W1 is split per node), SiLU -> smem tile;
//   3. 128x64x64 fp32 tile GEMM with W2 (resident in smem), bias+SiLU -> m, back to the smem tile;
//   4. segment sum of m over runs of equal row -> agg_m (one RED per (run, column); a run that
//      straddles a 32-edge slice or a tile adds its pieces with RED.ADD — no [E,64] tensor exists);
//   5. tile GEMM with Wc, SiLU, dot with w3 -> φ per edge; Δx·φ segment-summed into agg_x.
#include "common.cuh"

namespace degnn {

struct EdgeArgs {
    int64_t N, E;
    int A;
    unsigned flags;
    const int32_t* row;
    const int32_t* col;
    const float* ea;     // [E,A] in CSR order
    const float* x4;     // [N,4]
    const float* P;      // [N,64]
    const float* Q;      // [N,64]
    const float* w1r;    // [64]
    const float* w1e;    // [A][64]
    const float* w2;     // [64][64] k-major
    const float* b2;
    const float* wc;
    const float* bc;
    const float* w3;
    float* agg_m;        // [N,64] sums
    float* agg_x;        // [N,4] sums
};

constexpr int EDGE_SMEM_FLOATS = TILE_M * LDA        // activation tile
                                 + 2 * H * H         // W2, Wc
                                 + 5 * H             // b2, bc, w3, w1r (+ spare)
                                 + DISTEGNN_MAX_EDGE_ATTR * H   // w1e
                                 + TILE_M * DISTEGNN_MAX_EDGE_ATTR   // edge attrs of the tile
                                 + TILE_M * 4        // Δx
                                 + TILE_M            // φ
                                 + 2 * TILE_M;       // row, col (as int)
constexpr size_t EDGE_SMEM_BYTES = EDGE_SMEM_FLOATS * sizeof(float);

__global__ void __launch_bounds__(NTHREADS, 2) edge_layer_kernel(const EdgeArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* As = smem;
    float* W2s = As + TILE_M * LDA;
    float* Wcs = W2s + H * H;
    float* b2s = Wcs + H * H;
    float* bcs = b2s + H;
    float* w3s = bcs + H;
    float* w1rs = w3s + H;
    float* w1es = w1rs + 2 * H;
    float* eas = w1es + DISTEGNN_MAX_EDGE_ATTR * H;
    float* dxs = eas + TILE_M * DISTEGNN_MAX_EDGE_ATTR;
    float* phis = dxs + TILE_M * 4;
    int* srow = reinterpret_cast<int*>(phis + TILE_M);
    int* scol = srow + TILE_M;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    const int A = a.A;
    const bool normalize = a.flags & DISTEGNN_FLAG_NORMALIZE;
    const bool need_m = !(a.flags & DISTEGNN_FLAG_LAST);

    load_w64(W2s, a.w2, tid);
    load_w64(Wcs, a.wc, tid);
    if (tid < H) {
        b2s[tid] = a.b2[tid];
        bcs[tid] = a.bc[tid];
        w3s[tid] = a.w3[tid];
        w1rs[tid] = a.w1r[tid];
    }
    for (int i = tid; i < A * H; i += NTHREADS) w1es[i] = a.w1e[i];
    __syncthreads();

    const int64_t num_tiles = (a.E + TILE_M - 1) / TILE_M;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t e0 = tile * TILE_M;
        // ---- 1. edge metadata ----
        if (tid < TILE_M) {
            int64_t e = e0 + tid;
            bool ok = e < a.E;
            srow[tid] = ok ? __ldg(a.row + e) : -1;
            scol[tid] = ok ? __ldg(a.col + e) : 0;
        }
        for (int i = tid; i < TILE_M * A; i += NTHREADS) {
            int64_t g = e0 * A + i;
            eas[i] = (g < a.E * A) ? __ldg(a.ea + g) : 0.f;
        }
        __syncthreads();

        // ---- 2. gather + first layer: half-warp per edge, 16 edges per warp ----
        {
            const int l = lane & 15;
            const float4 wr4 = *reinterpret_cast<const float4*>(w1rs + 4 * l);
#pragma unroll 4
            for (int it = 0; it < 8; ++it) {
                const int el = 16 * warp + 2 * it + (lane >> 4);
                const int r = srow[el];
                float4 pre = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r >= 0) {
                    const int c = scol[el];
                    float4 p = ldg4(a.P + (size_t)r * H + 4 * l);
                    float4 q = ldg4(a.Q + (size_t)c * H + 4 * l);
                    float4 xi = ldg4(a.x4 + (size_t)r * 4);
                    float4 xj = ldg4(a.x4 + (size_t)c * 4);
                    float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
                    float radial = dx * dx + dy * dy + dz * dz;
                    if (normalize) {
                        float inv = 1.0f / (sqrtf(radial) + 1e-8f);
                        dx *= inv; dy *= inv; dz *= inv;
                    }
                    pre = fma4(radial, wr4, add4(p, q));
                    for (int k = 0; k < A; ++k)
                        pre = fma4(eas[el * A + k], *reinterpret_cast<const float4*>(w1es + k * H + 4 * l),
                                   pre);
                    pre = silu4(pre);
                    if (l == 0) *reinterpret_cast<float4*>(dxs + 4 * el) = make_float4(dx, dy, dz, 0.f);
                }
                *reinterpret_cast<float4*>(As + el * LDA + 4 * l) = pre;
            }
        }
        __syncthreads();

        // ---- 3. m = SiLU(W2·a1 + b2) ----
        float acc[8][4];
        zero_acc(acc);
        gemm_tile(acc, As, W2s, ty, tx);
        __syncthreads();   // everyone finished reading a1
        bias_silu_to_tile(acc, *reinterpret_cast<const float4*>(b2s + 4 * tx), As, ty, tx);
        __syncthreads();

        // ---- 4. Σ_j m_ij per destination row ----
        if (need_m) {
            const int c = tid & 63, q = tid >> 6;
            const int eb = 32 * q;
            int cur = srow[eb];
            float s = 0.f;
#pragma unroll 4
            for (int e = eb; e < eb + 32; ++e) {
                const int r = srow[e];
                if (r != cur) {
                    if (cur >= 0) atomicAdd(a.agg_m + (size_t)cur * H + c, s);
                    s = 0.f;
                    cur = r;
                }
                s += As[e * LDA + c];
            }
            if (cur >= 0) atomicAdd(a.agg_m + (size_t)cur * H + c, s);
        }

        // ---- 5. φ = w3·SiLU(Wc·m + bc); Σ_j Δx_ij·φ_ij ----
        zero_acc(acc);
        gemm_tile(acc, As, Wcs, ty, tx);
        head_dot_to_smem(acc, *reinterpret_cast<const float4*>(bcs + 4 * tx),
                         *reinterpret_cast<const float4*>(w3s + 4 * tx), phis, ty, tx);
        __syncthreads();
        if (tid < 12) {
            const int comp = tid % 3, q = tid / 3;
            const int eb = 32 * q;
            int cur = srow[eb];
            float s = 0.f;
            for (int e = eb; e < eb + 32; ++e) {
                const int r = srow[e];
                if (r != cur) {
                    if (cur >= 0) atomicAdd(a.agg_x + (size_t)cur * 4 + comp, s);
                    s = 0.f;
                    cur = r;
                }
                if (r >= 0) s = fmaf(dxs[4 * e + comp], phis[e], s);
            }
            if (cur >= 0) atomicAdd(a.agg_x + (size_t)cur * 4 + comp, s);
        }
        __syncthreads();   // tile buffers are rewritten by the next iteration
    }
}

}  // namespace degnn

extern "C" int distegnn_edge_layer_fwd_simt(int64_t n_nodes, int64_t n_edges, int A, int C, int Na,
                                       unsigned flags, const int32_t* row, const int32_t* col,
                                       const float* edge_attr_sorted, const float* x4, const float* P,
                                       const float* Q, const float* layer_params, float* agg_m,
                                       float* agg_x, void* stream) {
    using namespace degnn;
    if (int rc = check_dims(A, C, Na)) return rc;
    if (n_edges == 0) return DISTEGNN_OK;
    DEGNN_CHECK_ARG(n_nodes > 0 && n_edges > 0, "negative size");
    DEGNN_CHECK_ARG(row && col && x4 && P && Q && layer_params && agg_x, "null pointer");
    DEGNN_CHECK_ARG(A == 0 || edge_attr_sorted, "null edge_attr with edge_attr_nf > 0");
    DEGNN_CHECK_ARG((flags & DISTEGNN_FLAG_LAST) || agg_m, "null agg_m");
    Layout L = make_layout(A, C, Na);
    EdgeArgs a;
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

    static bool attr_set = false;
    if (!attr_set) {
        ensure_dynamic_smem((const void*)edge_layer_kernel, (int)EDGE_SMEM_BYTES);
        attr_set = true;
    }
    int64_t tiles = (n_edges + TILE_M - 1) / TILE_M;
    int64_t grid = (int64_t)sm_count() * 2;
    if (grid > tiles) grid = tiles;
    edge_layer_kernel<<<(unsigned)grid, NTHREADS, EDGE_SMEM_BYTES, (cudaStream_t)stream>>>(a);
    DEGNN_CHECK_LAUNCH();
    return DISTEGNN_OK;
}
