// Per-node stages: the embedding prologue, the node/coordinate update of one E_GCL_vel layer, and the
// tiny per-graph virtual-node update that runs after the all-reduce.  Replaces
//   FastEGNN.forward prologue                       models/FastEGNN.py:298-302
//   coord_model_vel (sum of the three terms), φ_v    :177-183
//   node_model                                       :203-217
//   coord_model_virtual / node_model_virtual (global halves), m_X   :199, :229-233, :258-264
// and produces the per-node operands of the NEXT layer's fused stages (P, Q, Hn — the per-node halves
// of the first edge-MLP / virtual-MLP layers, SURVEY §7 "W1 split").
#include "common.cuh"

namespace degnn {

// ---- block-level accumulation of (x', 1) per graph ---------------------------------------------
// tid < nvalid owns one node with value v (xyz, 1).  Single-graph tiles reduce in the CTA and add to
// the CTA-persistent accumulator accS[4]; mixed tiles fall back to global atomics.
__device__ __forceinline__ void accumulate_xsum(float4 v, bool valid, int g, bool single, float* accS,
                                                float* red /*[8][4]*/, float* vsum, int K, int tid) {
    if (single) {
        float s[4] = {valid ? v.x : 0.f, valid ? v.y : 0.f, valid ? v.z : 0.f, valid ? 1.f : 0.f};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s[j] += __shfl_xor_sync(FULL, s[j], o);
        }
        if ((tid & 31) == 0 && tid < TILE_M) {
#pragma unroll
            for (int j = 0; j < 4; ++j) red[(tid >> 5) * 4 + j] = s[j];
        }
    } else if (valid) {
        float* dst = vsum + (size_t)g * K;
        atomicAdd(dst + 0, v.x);
        atomicAdd(dst + 1, v.y);
        atomicAdd(dst + 2, v.z);
        atomicAdd(dst + 3, 1.0f);
    }
    __syncthreads();
    if (single && tid < 4) accS[tid] += red[tid] + red[4 + tid] + red[8 + tid] + red[12 + tid];
}

// Next-layer per-node operands from the h' tile sitting in As: P = W1a·h'+b1, Q = W1b·h', Hn = W1vh·h'.
__device__ __forceinline__ void emit_pqh(const float* As, float* Ws, const float* nw1a, const float* nb1,
                                         const float* nw1b, const float* nw1h, float* P, float* Q,
                                         float* Hn, int64_t n0, int nvalid, int tid, int ty, int tx) {
    const float* wsrc[3] = {nw1a, nw1b, nw1h};
    float* dsts[3] = {P, Q, Hn};
#pragma unroll 1
    for (int s = 0; s < 3; ++s) {
        __syncthreads();   // previous users of Ws are done
        load_w64(Ws, wsrc[s], tid);
        __syncthreads();
        float acc[8][4];
        zero_acc(acc);
        gemm_tile(acc, As, Ws, ty, tx);
        float4 b = (s == 0) ? ldg4(nb1 + 4 * tx) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = ty + 16 * i;
            if (r < nvalid)
                *reinterpret_cast<float4*>(dsts[s] + (size_t)(n0 + r) * H + 4 * tx) =
                    make_float4(acc[i][0] + b.x, acc[i][1] + b.y, acc[i][2] + b.z, acc[i][3] + b.w);
        }
    }
}

constexpr int NODE_SMEM_FLOATS = TILE_M * LDA + H * H + 3 * TILE_M + 16 + 4 + 4;
constexpr size_t NODE_SMEM_BYTES = NODE_SMEM_FLOATS * sizeof(float);

// =================================================================================================
// embedding prologue
// =================================================================================================
struct EmbedArgs {
    int64_t N;
    int B, F, K;
    const float* feat;       // [N,F]
    const float* loc;        // [N,3]
    const int64_t* batch64;  // [N]
    const float* wt;         // [F][64]
    const float* bias;       // [64]
    const float* nw1a; const float* nb1; const float* nw1b; const float* nw1h;
    float* h; float* x4; int32_t* batch32; float* P; float* Q; float* Hn; float* vsum;
};

__global__ void __launch_bounds__(NTHREADS, 2) embed_kernel(const EmbedArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* As = smem;
    float* Ws = As + TILE_M * LDA;
    float* scr = Ws + H * H;                 // [3*TILE_M] scratch (unused here)
    float* red = scr + 3 * TILE_M;           // [16]
    float* accS = red + 16;                  // [4]
    int* sg = reinterpret_cast<int*>(accS + 4);   // [2] first/last graph of the tile
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    if (tid < 4) accS[tid] = 0.f;
    int cur_graph = -1;
    __syncthreads();

    const int64_t num_tiles = (a.N + TILE_M - 1) / TILE_M;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t n0 = tile * TILE_M;
        const int nvalid = (int)min((int64_t)TILE_M, a.N - n0);
        // graph ids / coordinates / counts
        int g = -1;
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid < nvalid) {
            g = (int)a.batch64[n0 + tid];
            a.batch32[n0 + tid] = g;
            const float* p = a.loc + (size_t)(n0 + tid) * 3;
            xv = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
            *reinterpret_cast<float4*>(a.x4 + (size_t)(n0 + tid) * 4) = xv;
        }
        if (tid == 0) sg[0] = g;
        if (tid == nvalid - 1) sg[1] = g;
        __syncthreads();
        const bool single = sg[0] == sg[1];
        if (single && sg[0] != cur_graph) {
            if (cur_graph >= 0 && tid < 4) {
                atomicAdd(a.vsum + (size_t)cur_graph * a.K + tid, accS[tid]);
                accS[tid] = 0.f;
            }
            cur_graph = sg[0];
        }
        accumulate_xsum(xv, tid < nvalid, g, single, accS, red, a.vsum, a.K, tid);

        // h0 = feat·Wᵀ + b : thread per (row, 4 cols)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = ty + 16 * i;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nvalid) {
                v = ldg4(a.bias + 4 * tx);
                const float* f = a.feat + (size_t)(n0 + r) * a.F;
                for (int k = 0; k < a.F; ++k) v = fma4(__ldg(f + k), ldg4(a.wt + k * H + 4 * tx), v);
                *reinterpret_cast<float4*>(a.h + (size_t)(n0 + r) * H + 4 * tx) = v;
            }
            *reinterpret_cast<float4*>(As + r * LDA + 4 * tx) = v;
        }
        emit_pqh(As, Ws, a.nw1a, a.nb1, a.nw1b, a.nw1h, a.P, a.Q, a.Hn, n0, nvalid, tid, ty, tx);
        __syncthreads();
    }
    if (cur_graph >= 0 && tid < 4) atomicAdd(a.vsum + (size_t)cur_graph * a.K + tid, accS[tid]);
}

// =================================================================================================
// node update
// =================================================================================================
struct NodeArgs {
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

__global__ void __launch_bounds__(NTHREADS, 2) node_layer_kernel(const NodeArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* As = smem;
    float* Ws = As + TILE_M * LDA;
    float* phiv = Ws + H * H;                // [TILE_M]
    float* invdeg = phiv + TILE_M;           // [TILE_M]
    float* spare = invdeg + TILE_M;          // [TILE_M]
    float* red = spare + TILE_M;             // [16]
    float* accS = red + 16;                  // [4]
    int* sg = reinterpret_cast<int*>(accS + 4);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const bool last = a.flags & DISTEGNN_FLAG_LAST;
    if (tid < 4) accS[tid] = 0.f;
    int cur_graph = -1;
    __syncthreads();

    const int64_t num_tiles = (a.N + TILE_M - 1) / TILE_M;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t n0 = tile * TILE_M;
        const int nvalid = (int)min((int64_t)TILE_M, a.N - n0);
        int g = -1;
        if (tid < TILE_M) {
            float inv = 0.f;
            if (tid < nvalid) {
                g = __ldg(a.batch + n0 + tid);
                int deg = __ldg(a.rowptr + n0 + tid + 1) - __ldg(a.rowptr + n0 + tid);
                inv = 1.0f / (float)max(deg, 1);
            }
            invdeg[tid] = inv;
        }
        if (tid == 0) sg[0] = g;
        if (tid == nvalid - 1) sg[1] = g;
        // ---- S1: φ_v(h) from the OLD h (FastEGNN.py:183) ----
        load_a_tile(As, a.h + (size_t)n0 * H, nvalid, nullptr, tid);
        load_w64(Ws, a.lw, tid);
        __syncthreads();
        float acc[8][4];
        zero_acc(acc);
        gemm_tile(acc, As, Ws, ty, tx);
        head_dot_to_smem(acc, ldg4(a.lb + 4 * tx), ldg4(a.lw3 + 4 * tx), phiv, ty, tx);
        __syncthreads();

        // ---- coordinate update: thread per node ----
        const bool single = sg[0] == sg[1];
        if (single && sg[0] != cur_graph) {
            if (cur_graph >= 0 && tid < 4) {
                atomicAdd(a.vsum + (size_t)cur_graph * a.K + tid, accS[tid]);
                accS[tid] = 0.f;
            }
            cur_graph = sg[0];
        }
        float4 xn = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid < nvalid) {
            const size_t n = (size_t)(n0 + tid);
            float4 x = ldg4(a.x4 + n * 4);
            float4 ax = ldg4(a.agg_x + n * 4);
            float4 tv = ldg4(a.trans_v + n * 4);
            const float* v = a.vel + n * 3;
            const float pv = phiv[tid] + __ldg(a.lb3);
            const float id = invdeg[tid];
            xn.x = x.x + ax.x * id + tv.x + pv * __ldg(v);
            xn.y = x.y + ax.y * id + tv.y + pv * __ldg(v + 1);
            xn.z = x.z + ax.z * id + tv.z + pv * __ldg(v + 2);
            *reinterpret_cast<float4*>(a.x4_out + n * 4) = xn;
            if (a.loc_out) {
                a.loc_out[n * 3 + 0] = xn.x;
                a.loc_out[n * 3 + 1] = xn.y;
                a.loc_out[n * 3 + 2] = xn.z;
            }
        }
        accumulate_xsum(xn, tid < nvalid, g, single, accS, red, a.vsum, a.K, tid);
        if (last) {
            __syncthreads();
            continue;
        }

        // ---- node MLP layer 1: K = 64 (h) + 64 (agg) + 64 (agg_v) + Na ----
        zero_acc(acc);
        __syncthreads();
        load_w64(Ws, a.n1, tid);                               // As still holds h
        __syncthreads();
        gemm_tile(acc, As, Ws, ty, tx);
        __syncthreads();
        load_a_tile(As, a.agg_m + (size_t)n0 * H, nvalid, invdeg, tid);
        load_w64(Ws, a.n1 + H * H, tid);
        __syncthreads();
        gemm_tile(acc, As, Ws, ty, tx);
        __syncthreads();
        load_a_tile(As, a.agg_v + (size_t)n0 * H, nvalid, nullptr, tid);
        load_w64(Ws, a.n1 + 2 * H * H, tid);
        __syncthreads();
        gemm_tile(acc, As, Ws, ty, tx);
        if (a.Na > 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = ty + 16 * i;
                if (r < nvalid) {
                    const float* at = a.attr + (size_t)(n0 + r) * a.Na;
                    for (int k = 0; k < a.Na; ++k) {
                        float4 w = ldg4(a.n1 + (size_t)(3 * H + k) * H + 4 * tx);
                        float s = __ldg(at + k);
                        acc[i][0] = fmaf(s, w.x, acc[i][0]);
                        acc[i][1] = fmaf(s, w.y, acc[i][1]);
                        acc[i][2] = fmaf(s, w.z, acc[i][2]);
                        acc[i][3] = fmaf(s, w.w, acc[i][3]);
                    }
                }
            }
        }
        __syncthreads();
        bias_silu_to_tile(acc, ldg4(a.nb1 + 4 * tx), As, ty, tx);
        load_w64(Ws, a.n2, tid);
        __syncthreads();
        // ---- layer 2 + residual ----
        zero_acc(acc);
        gemm_tile(acc, As, Ws, ty, tx);
        __syncthreads();
        {
            const float4 b = ldg4(a.nb2 + 4 * tx);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = ty + 16 * i;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < nvalid) {
                    float4 h0 = ldg4(a.h + (size_t)(n0 + r) * H + 4 * tx);
                    v = make_float4(h0.x + acc[i][0] + b.x, h0.y + acc[i][1] + b.y,
                                    h0.z + acc[i][2] + b.z, h0.w + acc[i][3] + b.w);
                }
                *reinterpret_cast<float4*>(As + r * LDA + 4 * tx) = v;
            }
        }
        __syncthreads();
        // h' to HBM only after every thread has re-read the old h (h_out may alias h)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = ty + 16 * i;
            if (r < nvalid)
                *reinterpret_cast<float4*>(a.h_out + (size_t)(n0 + r) * H + 4 * tx) =
                    *reinterpret_cast<const float4*>(As + r * LDA + 4 * tx);
        }
        emit_pqh(As, Ws, a.nw1a, a.nxb1, a.nw1b, a.nw1h, a.P, a.Q, a.Hn, n0, nvalid, tid, ty, tx);
        __syncthreads();
    }
    if (cur_graph >= 0 && tid < 4) atomicAdd(a.vsum + (size_t)cur_graph * a.K + tid, accS[tid]);
}

}  // namespace degnn

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" int distegnn_embed_fwd_simt(int64_t n_nodes, int n_graphs, int F, int A, int C, int Na,
                                  const float* node_feat, const float* node_loc,
                                  const int64_t* data_batch, const float* emb_wt, const float* emb_b,
                                  const float* layer0_params, float* h, float* x4, int32_t* batch32,
                                  float* P, float* Q, float* Hn, float* vsum, void* stream) {
    using namespace degnn;
    if (int rc = check_dims(A, C, Na)) return rc;
    if (n_nodes == 0) return DISTEGNN_OK;
    DEGNN_CHECK_ARG(n_nodes > 0 && n_graphs > 0, "bad size");
    DEGNN_CHECK_ARG(F >= 1 && F <= DISTEGNN_MAX_NODE_FEAT, "node_feat_nf out of range");
    DEGNN_CHECK_ARG(node_feat && node_loc && data_batch && emb_wt && emb_b && layer0_params && h && x4 &&
                        batch32 && P && Q && Hn && vsum, "null pointer");
    Layout L = make_layout(A, C, Na);
    EmbedArgs a;
    a.N = n_nodes; a.B = n_graphs; a.F = F; a.K = 4 + 3 * C + H * C;
    a.feat = node_feat; a.loc = node_loc; a.batch64 = data_batch; a.wt = emb_wt; a.bias = emb_b;
    a.nw1a = layer0_params + L.off[DISTEGNN_P_E_W1A];
    a.nb1 = layer0_params + L.off[DISTEGNN_P_E_B1];
    a.nw1b = layer0_params + L.off[DISTEGNN_P_E_W1B];
    a.nw1h = layer0_params + L.off[DISTEGNN_P_V_W1H];
    a.h = h; a.x4 = x4; a.batch32 = batch32; a.P = P; a.Q = Q; a.Hn = Hn; a.vsum = vsum;
    ensure_dynamic_smem((const void*)embed_kernel, (int)NODE_SMEM_BYTES);
    int64_t tiles = (n_nodes + TILE_M - 1) / TILE_M;
    int64_t grid = (int64_t)sm_count() * 2;
    if (grid > tiles) grid = tiles;
    embed_kernel<<<(unsigned)grid, NTHREADS, NODE_SMEM_BYTES, (cudaStream_t)stream>>>(a);
    DEGNN_CHECK_LAUNCH();
    return DISTEGNN_OK;
}

extern "C" int distegnn_node_layer_fwd_simt(int64_t n_nodes, int n_graphs, int A, int C, int Na,
                                       unsigned flags, const int32_t* rowptr, const int32_t* batch32,
                                       const float* h, const float* x4, const float* node_vel,
                                       const float* node_attr, const float* agg_m, const float* agg_x,
                                       const float* agg_v, const float* trans_v,
                                       const float* layer_params, const float* next_layer_params,
                                       float* h_out, float* x4_out, float* P, float* Q, float* Hn,
                                       float* node_loc_out, float* vsum, void* stream) {
    using namespace degnn;
    if (int rc = check_dims(A, C, Na)) return rc;
    if (n_nodes == 0) return DISTEGNN_OK;
    const bool last = flags & DISTEGNN_FLAG_LAST;
    DEGNN_CHECK_ARG(n_nodes > 0 && n_graphs > 0, "bad size");
    DEGNN_CHECK_ARG(rowptr && batch32 && h && x4 && node_vel && agg_x && trans_v && layer_params &&
                        x4_out && vsum, "null pointer");
    DEGNN_CHECK_ARG(Na == 0 || last || node_attr, "null node_attr with node_attr_nf > 0");
    DEGNN_CHECK_ARG(last || (agg_m && agg_v && next_layer_params && h_out && P && Q && Hn),
                    "null pointer (non-last layer)");
    Layout L = make_layout(A, C, Na);
    NodeArgs a;
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
    a.h_out = h_out; a.x4_out = x4_out; a.P = P; a.Q = Q; a.Hn = Hn; a.loc_out = node_loc_out;
    a.vsum = vsum;
    ensure_dynamic_smem((const void*)node_layer_kernel, (int)NODE_SMEM_BYTES);
    int64_t tiles = (n_nodes + TILE_M - 1) / TILE_M;
    int64_t grid = (int64_t)sm_count() * 2;
    if (grid > tiles) grid = tiles;
    node_layer_kernel<<<(unsigned)grid, NTHREADS, NODE_SMEM_BYTES, (cudaStream_t)stream>>>(a);
    DEGNN_CHECK_LAUNCH();
    return DISTEGNN_OK;
}
