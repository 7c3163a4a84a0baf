/*
 * distegnn_b200.h — C ABI of libdistegnn_b200.so: the sm_100a implementation of the DistEGNN hot
 * path (FastEGNN per-layer equivariant message passing + the packed virtual-node statistics that are
 * all-reduced across graph partitions).
 *
 * The reference (GLAD-RUC/DistEGNN) is pure Python/PyTorch and has no FFI of its own; the entry
 * points below are what a binding for this path has to expose.  Each one cites the reference code it
 * replaces (paths relative to the reference repo root).  The only caller in this repo is
 * distegnn_b200/fast_egnn.py (ctypes); INTEGRATION.md shows the stub a maintainer of the reference
 * would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative DISTEGNN_E* code; distegnn_last_error()
 *     returns a thread-local human-readable message for the last failure;
 *   - all pointers are DEVICE pointers unless the name ends in _host; the caller owns every buffer
 *     (inputs, outputs, workspaces); the library never allocates, frees or synchronises — the one
 *     exception is the communicator (distegnn_comm_init / _destroy), which owns its peer-mapped segment;
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*); the call returns as soon as
 *     the work is enqueued; it is safe inside CUDA-graph capture;
 *   - fp32 everywhere, node ids int32 after distegnn_build_csr (int64 at the reference boundary);
 *   - hidden width H is fixed at 64 (every shipped config: config/ *.yaml `hidden_nf: 64`);
 *   - preconditions the kernels rely on and the host mirror validates (fast_egnn.py): data_batch is
 *     non-decreasing with ids in [0, n_graphs) (FastEGNN.py:298 takes B from data_batch[-1]+1 and PyG batches are
 *     sorted); edge ids lie in [0, n_nodes) (distegnn_build_csr returns DISTEGNN_EINVAL otherwise);
 *   - cross-check twins of these entry points (fp32-FMA, 3xTF32, alternative tcgen05 flavours) live in
 *     distegnn_b200_testing.h / libdistegnn_b200_testing.so and are not part of the product.
 *
 * Internal data layout (all row-major, contiguous)
 *   h      [N,64]   node features                x4     [N,4]  coordinates (xyz, w unused)
 *   P,Q    [N,64]   first edge-MLP layer split per node: P = W1[:, 0:64]·h + b1, Q = W1[:,64:128]·h
 *   Hn     [N,64]   first virtual-MLP layer node part:  W1v[:, 0:64]·h
 *   Xv     [B,3,C]  virtual coordinates (reference layout)
 *   Hv     [B,C,64] virtual features (reference layout is [B,64,C]; transposed once on the host)
 *   G      [B,C,64] per-graph/channel constant of the first virtual-MLP layer:
 *                   W1v[:,64:128]·Hv[b,:,c] + W1v[:,129:129+C]·m_X[b,:,c] + b1v
 *   vsum   [B,K]    packed per-graph partial sums that are all-reduced (SUM) once per layer,
 *                   K = 4 + 3C + 64C:  [0:3] Σ_i x_i   [3] node count   [4 : 4+3C] Σ_i ΔX_ic·φ_X as
 *                   [3][C]   [4+3C : K] Σ_i mv_ic as [C][64]
 */
#ifndef DISTEGNN_B200_H
#define DISTEGNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DISTEGNN_ABI_VERSION 2

#if defined(__GNUC__)
#define DISTEGNN_API __attribute__((visibility("default")))
#else
#define DISTEGNN_API
#endif

enum {
    DISTEGNN_OK = 0,
    DISTEGNN_EINVAL = -1,   /* bad argument (null pointer, unsupported size)              */
    DISTEGNN_ECUDA = -2,    /* a CUDA runtime call or kernel launch failed                 */
    DISTEGNN_EWORKSPACE = -3 /* caller-provided workspace too small                        */
};

/* limits of the compiled kernels */
#define DISTEGNN_HIDDEN 64
#define DISTEGNN_MAX_CHANNELS 16   /* virtual_channels C            */
#define DISTEGNN_MAX_EDGE_ATTR 8   /* edge_attr_nf A                */
#define DISTEGNN_MAX_NODE_ATTR 8   /* node_attr_nf Na               */
#define DISTEGNN_MAX_NODE_FEAT 16  /* node_feat_nf F                */

/* flags */
#define DISTEGNN_FLAG_NORMALIZE 1u  /* E_GCL_vel(normalize=True), FastEGNN.py:242-244                  */
#define DISTEGNN_FLAG_LAST 2u       /* last layer: h'/Hv' are dead (FastEGNN.py:307) — skip them        */
#define DISTEGNN_FLAG_INIT 4u       /* virtual_update before layer 0: no X / Hv update, only x̄, m_X, G */
#define DISTEGNN_FLAG_ZERO_VSUM 8u  /* virtual_update: leave vsum zeroed (ready for the next layer's accumulation)  */
#define DISTEGNN_FLAG_ZERO_AGG 16u  /* node_layer: leave agg_m / agg_x zeroed (ready for the next edge stage)       */

DISTEGNN_API int distegnn_abi_version(void);
DISTEGNN_API const char *distegnn_last_error(void);

/* ---- per-layer parameter block -------------------------------------------------------------------
 * One flat fp32 buffer per E_GCL_vel layer; every matrix is stored k-major ([in][out], i.e. the
 * transpose of nn.Linear.weight) so a thread reads 4 consecutive outputs with one 16-byte load.
 * Field order / source tensors (state_dict keys under gcl_<i>.):
 */
enum {
    DISTEGNN_P_E_W1A = 0,  /* [64][64]  edge_mlp.0.weight[:, 0:64]^T   (h[row])            FastEGNN.py:69-74 */
    DISTEGNN_P_E_W1B,      /* [64][64]  edge_mlp.0.weight[:, 64:128]^T (h[col])                              */
    DISTEGNN_P_E_W1R,      /* [64]      edge_mlp.0.weight[:, 128]      (radial)                              */
    DISTEGNN_P_E_W1E,      /* [A][64]   edge_mlp.0.weight[:, 129:129+A]^T (edge_attr)                        */
    DISTEGNN_P_E_B1,       /* [64]      edge_mlp.0.bias                                                      */
    DISTEGNN_P_E_W2,       /* [64][64]  edge_mlp.2.weight^T                                                  */
    DISTEGNN_P_E_B2,       /* [64]                                                                            */
    DISTEGNN_P_E_WC,       /* [64][64]  coord_mlp_r.0.weight^T                                 :96-110       */
    DISTEGNN_P_E_BC,       /* [64]                                                                            */
    DISTEGNN_P_E_W3,       /* [64]      coord_mlp_r.2.weight[0]                                               */
    DISTEGNN_P_V_W1H,      /* [64][64]  edge_mlp_virtual.0.weight[:, 0:64]^T (h)                :76-81        */
    DISTEGNN_P_V_W1V,      /* [64][64]  edge_mlp_virtual.0.weight[:, 64:128]^T (Hv)                           */
    DISTEGNN_P_V_W1R,      /* [64]      edge_mlp_virtual.0.weight[:, 128] (‖ΔX‖)                              */
    DISTEGNN_P_V_W1M,      /* [C][64]   edge_mlp_virtual.0.weight[:, 129:129+C]^T (m_X)                       */
    DISTEGNN_P_V_B1,       /* [64]                                                                            */
    DISTEGNN_P_V_W2,       /* [64][64]  edge_mlp_virtual.2.weight^T                                           */
    DISTEGNN_P_V_B2,       /* [64]                                                                            */
    DISTEGNN_P_V_WXV,      /* [64][64]  coord_mlp_r_virtual.0.weight^T                          :111          */
    DISTEGNN_P_V_BXV,      /* [64]                                                                            */
    DISTEGNN_P_V_W3XV,     /* [64]      coord_mlp_r_virtual.2.weight[0]                                       */
    DISTEGNN_P_V_WX,       /* [64][64]  coord_mlp_v_virtual.0.weight^T                          :112          */
    DISTEGNN_P_V_BX,       /* [64]                                                                            */
    DISTEGNN_P_V_W3X,      /* [64]      coord_mlp_v_virtual.2.weight[0]                                       */
    DISTEGNN_P_L_W,        /* [64][64]  coord_mlp_vel.0.weight^T                                :115-119      */
    DISTEGNN_P_L_B,        /* [64]                                                                            */
    DISTEGNN_P_L_W3,       /* [64]      coord_mlp_vel.2.weight[0]                                             */
    DISTEGNN_P_L_B3,       /* [4]       coord_mlp_vel.2.bias (1 value, padded)                                */
    DISTEGNN_P_N_W1,       /* [192+Na][64] node_mlp.0.weight^T  (h | agg | agg_v | node_attr)   :130-135      */
    DISTEGNN_P_N_B1,       /* [64]                                                                            */
    DISTEGNN_P_N_W2,       /* [64][64]  node_mlp.2.weight^T                                                   */
    DISTEGNN_P_N_B2,       /* [64]                                                                            */
    DISTEGNN_P_M_W1,       /* [128][64] node_mlp_virtual.0.weight^T (Hv | agg)                  :137-141      */
    DISTEGNN_P_M_B1,       /* [64]                                                                            */
    DISTEGNN_P_M_W2,       /* [64][64]  node_mlp_virtual.2.weight^T                                           */
    DISTEGNN_P_M_B2,       /* [64]                                                                            */
    DISTEGNN_P_NUM_FIELDS
};

/* Fills offsets_host[DISTEGNN_P_NUM_FIELDS] (in floats, each a multiple of 4) and *total_floats for a
 * layer with edge_attr_nf=A, virtual_channels=C, node_attr_nf=Na.  Host-only, no CUDA call. */
DISTEGNN_API int distegnn_param_layout(int A, int C, int Na, int64_t *offsets_host, int64_t *total_floats_host);

/* ---- graph preprocessing (cached per edge_index by the caller) -----------------------------------
 * Replaces the implicit "scatter by edge_index[0]" of unsorted_segment_sum/mean
 * (models/FastEGNN.py:322-337, twins models/basic.py:50-66): the COO list [2,E] int64 (row =
 * edge_index[0] = aggregation destination, col = edge_index[1] = neighbour; FastEGNN.py:238,250) is
 * stably sorted by row into int32 CSR.  perm[e'] = original position of sorted edge e' (to permute
 * edge_attr).  Self loops, duplicate edges and isolated nodes are legal (equivariant_test.py:26-27).
 * Ids outside [0, n_nodes) — on which the reference dies with a device-side index assert — are counted into the device
 * counter *n_invalid (may be NULL) and clamped, so that nothing indexes out of bounds; the caller reads the counter once
 * per graph (the build is cached) and raises.
 */
DISTEGNN_API int distegnn_csr_workspace_bytes(int64_t n_nodes, int64_t n_edges, int64_t *bytes_host);
DISTEGNN_API int distegnn_build_csr(const int64_t *edge_index, int64_t n_nodes, int64_t n_edges,
                       int32_t *rowptr /*[N+1]*/, int32_t *row /*[E]*/, int32_t *col /*[E]*/,
                       int32_t *perm /*[E]*/, void *workspace, int64_t workspace_bytes,
                       int32_t *n_invalid /*[1], device, may be NULL*/, void *stream);

/* dst[i,:] = src[perm[i],:] for i < n_rows, rows of `width` floats (edge_attr into CSR order). */
DISTEGNN_API int distegnn_gather_rows(const float *src, const int32_t *perm, int64_t n_rows, int width, float *dst,
                         void *stream);

/* ---- embedding + per-forward setup ---------------------------------------------------------------
 * FastEGNN.forward prologue (FastEGNN.py:298-302): h0 = embedding_in(node_feat); also converts
 * node_loc [N,3] → x4, data_batch int64 → batch32, computes P/Q/Hn of layer 0 from `layer0_params`,
 * and accumulates Σx and the node count of every graph into vsum[:,0:4] (caller zeroes vsum).
 * emb_wt is embedding_in.weight^T [F][64], emb_b [64].
 * data_batch must be non-decreasing with ids in [0, n_graphs): the per-graph reductions of every later stage rely on
 * it.  Violations are counted into the device counter *n_invalid (may be NULL; the caller zeroes it) and the ids clamped.
 */
DISTEGNN_API int distegnn_embed_fwd(int64_t n_nodes, int n_graphs, int F, int A, int C, int Na,
                       const float *node_feat, const float *node_loc, const int64_t *data_batch,
                       const float *emb_wt, const float *emb_b, const float *layer0_params,
                       float *h, float *x4, int32_t *batch32, float *P, float *Q, float *Hn,
                       float *vsum, int32_t *n_invalid, void *stream);

/* ---- real↔real edge stage ------------------------------------------------------------------------
 * coord2radial + edge_model + the edge part of coord_model_vel + the edge part of node_model
 * (FastEGNN.py:237-246, 144-150, 169-177, 206): for every CSR edge (i=row, j=col)
 *   Δx = x_i − x_j, r = ‖Δx‖² (Δx /= sqrt(r)+1e-8 if NORMALIZE)
 *   m  = SiLU(W2·SiLU(P_i + Q_j + w_r·r + W_e·a_ij) + b2),  φ = w3·SiLU(Wc·m + bc)
 *   agg_m[i] += m   (skipped with FLAG_LAST),   agg_x[i].xyz += Δx·φ
 * Sums, not means: the division by max(deg,1) happens in distegnn_node_layer_fwd.  The caller zeroes
 * agg_m [N,64] and agg_x [N,4] before the call.  For graphs built on the device without a host round trip
 * (distegnn_radius_graph_csr with a capacity) n_edges is the CAPACITY of row/col/edge_attr and n_edges_dev points to the
 * true count, which the kernel reads itself.
 */
DISTEGNN_API int distegnn_edge_layer_fwd(int64_t n_nodes, int64_t n_edges, int A, int C, int Na, unsigned flags,
                            const int32_t *row, const int32_t *col, const float *edge_attr_sorted,
                            const float *x4, const float *P, const float *Q,
                            const float *layer_params, float *agg_m, float *agg_x,
                            const int32_t *n_edges_dev /*NULL, or the edge count on the device (<= n_edges)*/,
                            void *stream);

/* Backward of distegnn_edge_layer_fwd (in the reference: autograd through models/FastEGNN.py:144-150,
 * 169-177, 206, 237-246, 322-337).  Nothing of size [E,.] is kept from the forward pass: every 128-edge tile is
 * recomputed.  Inputs: the forward inputs plus g_agg_m [N,64] (gradient w.r.t. the SUM agg_m; may be NULL with
 * FLAG_LAST) and g_agg_x [N,4] (w.r.t. the SUM agg_x).  Outputs are ACCUMULATED (+=): g_P, g_Q [N,64], g_x4 [N,4]
 * (both edge endpoints; the normalisation norm is detached as in :243) and g_layer_params, a buffer with the layout
 * of the parameter block (fields E_W1R, E_W1E, E_W2, E_B2, E_WC, E_BC, E_W3 are written). */
DISTEGNN_API int distegnn_edge_layer_bwd(int64_t n_nodes, int64_t n_edges, int A, int C, int Na, unsigned flags,
                                         const int32_t* row, const int32_t* col, const float* edge_attr_sorted,
                                         const float* x4, const float* P, const float* Q, const float* layer_params,
                                         const float* g_agg_m, const float* g_agg_x, float* g_P, float* g_Q,
                                         float* g_x4, float* g_layer_params, const int32_t* n_edges_dev, void* stream);

/* Backward of distegnn_virtual_layer_fwd (in the reference: autograd through models/FastEGNN.py:154-163,
 * 180, 191-193, 207, 220-223, 252-253).  Rows are recomputed tile by tile; the six row-wise tile GEMMs run on tcgen05.
 * `weight_images` (96 KB, device): the stage's three 64x64 matrices and their transposes as fp16 hi/lo images in the
 * shared-memory operand layout, written by distegnn_virtual_bwd_prepare(layer_params) — once per layer and call; the
 * kernel streams them through shared memory with TMA.  Upstream gradients: g_agg_v [N,64] (NULL with FLAG_LAST), g_trans_v
 * [N,4], g_vsum [B,K] (entries [4:] are read; already summed over the partitions).  g_Hn [N,64] and g_xv [N,4] are WRITTEN;
 * g_G [B,C,64], g_Xv [B,3,C] and the parameter gradients (V_W1R, V_W2, V_B2, V_WXV, V_BXV, V_W3XV, V_WX, V_BX, V_W3X of a
 * parameter-layout buffer) are accumulated. */
DISTEGNN_API int distegnn_virtual_bwd_prepare(int A, int C, int Na, const float* layer_params, void* weight_images,
                                              void* stream);
DISTEGNN_API int distegnn_virtual_layer_bwd(int64_t n_nodes, int n_graphs, int A, int C, int Na, unsigned flags,
                                            const int32_t* batch32, const float* x4, const float* Hn, const float* Xv,
                                            const float* G, const float* layer_params, const void* weight_images,
                                            const float* g_agg_v, const float* g_trans_v, const float* g_vsum,
                                            float* g_Hn, float* g_xv, float* g_G, float* g_Xv, float* g_layer_params,
                                            void* stream);

/* On-device radius graph: replaces the host-side `radius_graph(pos_i, r=radius, max_num_neighbors=N)`
 * + `edge_attr = |dx|` of the reference's partitioners (datasets/distribute_graphs.py:43-44; PyG / torch_cluster).
 * Uniform-grid cell list; the caller sorts the nodes' cell keys (key = graph*ncell + (ix*ny + iy)*nz + iz, cell size >=
 * radius, ix = (int)((x - origin_x) * (1/cell)) clamped to the grid) and passes `order` (node ids in key order) and the
 * dense table cell_start[n_graphs*ncell + 1] (first position of every key).  origin_host[3] / dims_host[3] are HOST
 * arrays.  Two phases because the edge count is only known after the first:
 *   distegnn_radius_count -> deg[i] = number of j (same graph, j != i unless loop) with |x_i - x_j| < radius
 *   caller: rowptr = exclusive prefix sum of deg (int64 [N+1]), allocates E = rowptr[N] entries
 *   distegnn_radius_fill  -> row[e] = i, col[e] = j for e in [rowptr[i], rowptr[i+1]), dist[e] = |x_i - x_j| (dist may be
 *                            NULL): edges grouped by destination row, rows ascending. */
DISTEGNN_API int distegnn_radius_count(int64_t n_nodes, const float* x4, const int32_t* batch32, const int32_t* order,
                                       const int64_t* cell_start, const float* origin_host, float cell_size,
                                       const int32_t* dims_host, float radius, int loop, int32_t* deg, void* stream);
DISTEGNN_API int distegnn_radius_fill(int64_t n_nodes, const float* x4, const int32_t* batch32, const int32_t* order,
                                      const int64_t* cell_start, const float* origin_host, float cell_size,
                                      const int32_t* dims_host, float radius, int loop, const int64_t* rowptr,
                                      int32_t* row, int32_t* col, float* dist, void* stream);

/* On-device radius graph, CSR out, in ONE call and without a host round trip (csrc/radius_csr.cu): the
 * reference-boundary tensors in (pos [N,3] fp32, data_batch int64 [N] sorted, may be NULL for one graph), int32 CSR by
 * destination out (rowptr [N+1], row / col [capacity]) plus edge_attr [capacity, edge_attr_nf] = the edge length in every
 * column (datasets/distribute_graphs.py:43-44).  Pairs with |x_i - x_j| < radius (strict, as torch_cluster), j != i unless
 * `loop`, same graph only.  Bounding box, grid sizing (cell >= radius, grown until graphs x cells <= table_cells), cell
 * keys, sort, counts and prefix sums all happen on the device; info [4] (device): [0] edges found, [1] 1 if that exceeds
 * `capacity` (then only rowptr is complete), [2] cells used.  capacity = 0 runs the count only (row/col may be NULL). */
DISTEGNN_API int distegnn_radius_csr_workspace_bytes(int64_t n_nodes, int64_t table_cells, int64_t *bytes_host);
DISTEGNN_API int distegnn_radius_graph_csr(int64_t n_nodes, int n_graphs, const float *pos, const int64_t *data_batch,
                                           float radius, int loop, int edge_attr_nf, int64_t capacity,
                                           int64_t table_cells, int32_t *rowptr, int32_t *row, int32_t *col,
                                           float *edge_attr, int32_t *info, void *workspace, int64_t workspace_bytes,
                                           void *stream);

/* Lloyd iterations of the k-means node partitioner (datasets/distribute_graphs.py:118-143, 188-198: sklearn KMeans on the
 * host) with sklearn's stopping rules evaluated on the device: `iters` iterations are enqueued; once no label changes, or
 * the squared centre shift is <= tol (then after one more assignment pass), the remaining ones are no-ops.  centers [K,3]
 * in/out (seeded by the caller, k-means++), labels int32 [N] in/out (−1 initially), sums float64 [K,4] zeroed once by the
 * caller, state int32 [4] zeroed once: [0] 0 running / 1 final pass pending / 2 converged, [1] iterations done. */
DISTEGNN_API int distegnn_kmeans_lloyd(int64_t n_nodes, int n_clusters, const float *pos, float *centers,
                                       int32_t *labels, double *sums, int32_t *state, float tol, int iters, void *stream);

/* ---- real↔virtual stage --------------------------------------------------------------------------
 * Virtual geometry + edge_mode_virtual + the virtual parts of coord_model_vel, coord_model_virtual,
 * node_model and node_model_virtual (FastEGNN.py:252-253, 154-163, 180, 191-193, 207, 220-223):
 * for every node i (graph b) and channel c
 *   ΔX = Xv[b,:,c] − x_i,  mv = SiLU(W2v·SiLU(Hn_i + G[b,c] + w_vr·‖ΔX‖) + b2v)
 *   trans_v[i] = mean_c(−ΔX·φ_xv(mv)),  agg_v[i] = mean_c mv                       (per node)
 *   vsum[b, 4:4+3C] += ΔX·φ_X(mv),      vsum[b, 4+3C:] += mv                       (per graph)
 * With FLAG_LAST agg_v and the Σmv block are skipped.  Caller zeroes vsum.
 */
DISTEGNN_API int distegnn_virtual_layer_fwd(int64_t n_nodes, int n_graphs, int A, int C, int Na, unsigned flags,
                               const int32_t *batch32, const float *x4, const float *Hn,
                               const float *Xv, const float *G, const float *layer_params,
                               float *agg_v, float *trans_v /*[N,4]*/, float *vsum, void *stream);

/* ---- node update ---------------------------------------------------------------------------------
 * The rest of coord_model_vel and node_model (FastEGNN.py:177-183, 203-217):
 *   x' = x + agg_x/max(deg,1) + trans_v + φ_v(h)·v
 *   h' = h + W2n·SiLU(W1n·[h; agg_m/max(deg,1); agg_v; node_attr] + b1n) + b2n
 * plus P/Q/Hn of the NEXT layer from next_layer_params, and vsum[b,0:4] += (x', 1).
 * With FLAG_LAST only x' is produced and additionally written as [N,3] to node_loc_out (the model
 * output); h_out/P/Q/Hn/next_layer_params may be null.  h_out may alias h, x4_out may alias x4.
 * With FLAG_ZERO_AGG the kernel clears agg_m and agg_x after consuming them (they are written although declared
 * const), so the caller needs no memset before the next edge stage.
 */
DISTEGNN_API int distegnn_node_layer_fwd(int64_t n_nodes, int n_graphs, int A, int C, int Na, unsigned flags,
                            const int32_t *rowptr, const int32_t *batch32, const float *h,
                            const float *x4, const float *node_vel, const float *node_attr,
                            const float *agg_m, const float *agg_x, const float *agg_v,
                            const float *trans_v, const float *layer_params,
                            const float *next_layer_params, float *h_out, float *x4_out, float *P,
                            float *Q, float *Hn, float *node_loc_out, float *vsum, void *stream);

/* Backward of distegnn_node_layer_fwd and of the embedding prologue (in the reference: autograd through
 * models/FastEGNN.py:177-183, 203-217, 302) — fp32 FMA tile GEMMs, everything recomputed from N-sized tensors.
 * Upstream: g_x_out [N,3] (w.r.t. x'), g_vsum [B,K] (optional; its [b,0:3] entries add to g_x' of the nodes of graph b —
 * x' feeds the next layer's Σ x'), g_h_out / g_P / g_Q / g_Hn [N,64] (w.r.t. h' and the next layer's projections; NULL with
 * FLAG_LAST).  WRITTEN: g_h [N,64], g_x [N,3], g_agg_x / g_trans_v [N,4], g_agg_m / g_agg_v [N,64] (not with FLAG_LAST).
 * ACCUMULATED: fields L_W, L_B, L_W3, L_B3, N_W1, N_B1, N_W2, N_B2 of g_layer_params and E_W1A, E_B1, E_W1B, V_W1H of
 * g_next_layer_params (parameter-layout buffers).  distegnn_embed_bwd: h0 = the forward's h of layer 0; accumulates
 * g_emb_wt [F][64], g_emb_b [64] and the same four projection fields of layer 0's gradient block. */
DISTEGNN_API int distegnn_node_layer_bwd(int64_t n_nodes, int A, int C, int Na, unsigned flags, const int32_t *rowptr,
                                         const float *h, const float *node_vel, const float *node_attr,
                                         const float *agg_m, const float *agg_v, const float *layer_params,
                                         const float *next_layer_params, const float *g_x_out, const float *g_vsum,
                                         const int32_t *batch32, const float *g_h_out, const float *g_P, const float *g_Q,
                                         const float *g_Hn, float *g_h, float *g_x, float *g_agg_x, float *g_trans_v,
                                         float *g_agg_m, float *g_agg_v, float *g_layer_params,
                                         float *g_next_layer_params, void *stream);
DISTEGNN_API int distegnn_embed_bwd(int64_t n_nodes, int F, int A, int C, int Na, const float *node_feat, const float *h0,
                                    const float *layer0_params, const float *g_h, const float *g_P, const float *g_Q,
                                    const float *g_Hn, float *g_emb_wt, float *g_emb_b, float *g_layer0_params,
                                    void *stream);

/* ---- virtual-node sync: packed SUM all-reduce over NVLink peer memory ------------------------------------------------
 * Replaces weighted_average_reduce / _AllReduce (models/FastEGNN.py:10-43, 310-319; call sites :195-197, 225-227,
 * 259-261 — six NCCL calls behind host syncs per layer) with ONE exchange of the packed statistics vsum [B,K] per layer.
 * One process per GPU; every rank owns a segment of device memory that all peers map through CUDA IPC.  A call pushes the
 * rank's values into every peer's segment, raises a per-slot flag, waits for the peers' flags and sums the `world`
 * contributions in RANK ORDER, so the result is bit-identical on all ranks (the property the reference relies on NCCL
 * for, FastEGNN.py:29-31) and independent of arrival order.  No host synchronisation, safe under CUDA-graph capture (the
 * per-slot epoch lives in device memory).  All ranks must issue the same sequence of calls with the same sizes.
 *
 *   distegnn_comm_init     allocate + clear this rank's segment for calls of at most max_slots x slot_floats floats (for
 *                          the model: max_slots >= B graphs, slot_floats >= K); returns the opaque communicator and this
 *                          rank's IPC handle (distegnn_comm_handle_bytes() bytes, host memory).  Synchronises the device.
 *   (caller)               all-gathers the handles over any host transport (the mirror uses torch.distributed)
 *   distegnn_comm_connect  all_handles_host = world handles in rank order; maps the peers' segments
 *   distegnn_allreduce_packed  in-place SUM of buf[0:count] over the ranks, enqueued on `stream`
 *   distegnn_comm_status   *status_host != 0 if a wait timed out (a peer never arrived; default 10 s, see _set_timeout_ms)
 *   distegnn_comm_disconnect  unmap the peers' segments (own segment stays: peers may still map it)
 *   distegnn_comm_destroy  unmap + free.  Teardown across ranks: no calls in flight -> every rank disconnects -> host
 *                          barrier -> every rank destroys (an exported allocation must outlive its importers' mappings)
 * The fused form — all-reduce of vsum[b,:] followed by the virtual-node update in the same kernel — is
 * distegnn_virtual_update_fwd with a non-null `comm`.
 */
DISTEGNN_API int distegnn_comm_handle_bytes(void);
DISTEGNN_API int distegnn_comm_init(int rank, int world, int max_slots, int slot_floats, void **comm_out,
                                    void *handle_out_host);
DISTEGNN_API int distegnn_comm_connect(void *comm, const void *all_handles_host);
DISTEGNN_API int distegnn_comm_set_timeout_ms(void *comm, int64_t milliseconds);
DISTEGNN_API int distegnn_comm_status(void *comm, int *status_host);
DISTEGNN_API int distegnn_comm_disconnect(void *comm);
DISTEGNN_API int distegnn_comm_destroy(void *comm);
DISTEGNN_API int distegnn_allreduce_packed(void *comm, float *buf, int64_t count, void *stream);

/* ---- virtual-node update, fused with the all-reduce of vsum ---------------------------------------------------------
 * The global halves of coord_model_virtual / node_model_virtual and the next layer's m_X
 * (FastEGNN.py:199, 229-233, 258-264) from the *summed* statistics (weighted_average_reduce,
 * FastEGNN.py:310-319, is Σ_r n_r·mean_r / Σ_r n_r = Σ_r sum_r / Σ_r n_r).  One CTA per graph:
 *   [comm != NULL]  vsum[b,:] := Σ over the partitions (the exchange described above, slot = graph)
 *   n = max(vsum[b,3],1);  Xv += vsum[b,4:4+3C]/n;  Hv += MLP_hv([Hv; vsum[b,4+3C:]/n])
 *   x̄ = vsum[b,0:3]/n;  m_X = (Xv−x̄)ᵀ(Xv−x̄);  G_next = W1v_V·Hv + W1v_M·m_X + b1v (next layer's)
 * FLAG_INIT: skip the Xv/Hv updates (before layer 0); if init_loc_mean [B,3] / init_hv0 [C,64] are given, Xv / Hv are
 * first initialised from them (FastEGNN.py:299-300) instead of being read.  FLAG_LAST: only Xv is updated.
 * FLAG_ZERO_VSUM: vsum is left zeroed for the next layer's accumulation; otherwise it holds the summed statistics on
 * return (the training path keeps them for the backward pass).
 * layer_params: this layer's block (node_mlp_virtual); next_layer_params: the block whose virtual MLP
 * consumes G (null with FLAG_LAST).  With FLAG_INIT pass layer 0's block as next_layer_params.
 */
DISTEGNN_API int distegnn_virtual_update_fwd(int n_graphs, int A, int C, int Na, unsigned flags, float *vsum,
                                float *Xv, float *Hv, const float *layer_params,
                                const float *next_layer_params, float *G, const float *init_loc_mean,
                                const float *init_hv0, void *comm, void *stream);

/* Backward of distegnn_virtual_update_fwd (per graph; in the reference: autograd through models/FastEGNN.py:193-199,
 * 222-234, 258-264).  vsum = the SUMMED statistics the forward consumed, Xv / Hv = the forward's inputs (with FLAG_INIT: the
 * initial loc_mean / virtual_node_feat broadcasts).  Upstream g_Xn [B,3,C], g_Hn, g_G [B,C,64] (NULL = zero).  WRITTEN:
 * g_vsum [B,K] (entry 3, the node count, gets 0), g_Xv [B,3,C], g_Hv [B,C,64] (not with FLAG_LAST).  ACCUMULATED: M_W1, M_B1,
 * M_W2, M_B2 of g_layer_params (regular layers) and V_W1V, V_W1M, V_B1 of g_next_layer_params. */
DISTEGNN_API int distegnn_virtual_update_bwd(int n_graphs, int A, int C, int Na, unsigned flags, const float *vsum,
                                             const float *Xv, const float *Hv, const float *layer_params,
                                             const float *next_layer_params, const float *g_Xn, const float *g_Hn,
                                             const float *g_G, float *g_vsum, float *g_Xv, float *g_Hv,
                                             float *g_layer_params, float *g_next_layer_params, void *stream);

/* ---- loss side of the training step -------------------------------------------------------------------
 * Replaces utils/train.py:98-147: node-count weighted MSE (:98-110), the MMD regulariser between the virtual
 * coordinates and S = samples·C sampled target positions per graph (:119-147, kernel k(x,y) = exp(−‖x−y‖₂/(2σ²)), :11-14)
 * and the per-step scalar collectives (:104, :109 and the loc_mean all_gather check :52-61), folded into ONE packed SUM
 * all-reduce issued by the caller between the two calls (distegnn_allreduce_packed or any SUM all-reduce).
 *   packed [2 + world·3B]  (zeroed by the caller)  [0] n_r, [1] n_r·MSE_r, [2 + r·3B ...] rank r's loc_mean
 *   acc    [3]             (zeroed by the caller)  Σ_b l_vv, Σ_b l_rv, n_r·MSE_r (local copies for the finalize)
 *   graph_ptr [B+1] int64: first node of every graph (data_batch is sorted); samples [B,S] int32: node indices LOCAL to
 *   the graph drawn by the caller (the reference uses torch.randperm(num_node)[:S] per graph; −1 pads graphs with fewer
 *   than S nodes — the reference still divides by S)
 * distegnn_loss_finalize (after the all-reduce): out[0] = loss to back-propagate = world·n_r/Σn·(MSE_r + weight·MMD_r) /
 * accumulation_steps, out[1] = logged loss Σ_r n_r/Σn·MSE_r, out[2] = MMD_r, out[3] = max |loc_mean_r − loc_mean_0|, and
 * the gradients of out[0]: g_pred [N,3], g_Xv [B,3,C] (cdist's convention: zero gradient at coincident points). */
DISTEGNN_API int distegnn_loss_packed_floats(int n_graphs, int world);
DISTEGNN_API int distegnn_loss_partials(int64_t n_nodes, int n_graphs, int C, int S, int world, int rank, float sigma,
                                        const float *pred, const float *target, const float *Xv, const float *loc_mean,
                                        const int64_t *graph_ptr, const int32_t *samples, float *acc, float *packed,
                                        float *gV_raw, void *stream);
DISTEGNN_API int distegnn_loss_finalize(int64_t n_nodes, int n_graphs, int C, int S, int world, int rank, float sigma,
                                        float weight, int accumulation_steps, const float *pred, const float *target,
                                        const float *loc_mean, const float *acc, const float *packed, const float *gV_raw,
                                        float *g_pred, float *g_Xv, float *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DISTEGNN_B200_H */
