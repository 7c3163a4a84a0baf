/*
 * distegnn_b200_testing.h — cross-check twins of the production entry points (libdistegnn_b200_testing.so).
 *
 * NOT part of the product path: FastEGNN.forward never calls these.  They are earlier / alternative implementations of
 * the same stages (fp32-FMA on the CUDA cores, 3xTF32, thread-per-row and column-split tcgen05 flavours) kept as
 * independent implementations for parity tests at sizes the CPU oracle cannot reach and for A/B timing
 * (tests/twin_backend.py is the only caller).  Contracts are those of the production symbols in distegnn_b200.h.
 */
#ifndef DISTEGNN_B200_TESTING_H
#define DISTEGNN_B200_TESTING_H

#include "distegnn_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* fp32-FMA twin of distegnn_embed_fwd (cross-check only). */
DISTEGNN_API int distegnn_embed_fwd_simt(int64_t n_nodes, int n_graphs, int F, int A, int C, int Na,
                                         const float *node_feat, const float *node_loc, const int64_t *data_batch,
                                         const float *emb_wt, const float *emb_b, const float *layer0_params,
                                         float *h, float *x4, int32_t *batch32, float *P, float *Q, float *Hn,
                                         float *vsum, void *stream);

/* Same contract as distegnn_virtual_layer_fwd: the column-split flavour (two threads per row, 32 warps per SM;
 * csrc/virtual_layer_cs.cu) — measured 2 % slower than the production thread-per-row kernel; kept as a twin. */
DISTEGNN_API int distegnn_virtual_layer_fwd_cs(int64_t n_nodes, int n_graphs, int A, int C, int Na, unsigned flags,
                                                const int32_t* batch32, const float* x4, const float* Hn,
                                                const float* Xv, const float* G, const float* layer_params,
                                                float* agg_v, float* trans_v, float* vsum, void* stream);

/* Same outputs with every tile GEMM as fp32 FMA on the CUDA cores (csrc/virtual_layer_bwd.cu; the first backward kernel,
 * kept as a twin).  wT = the matrices V_W2, V_WXV, V_WX TRANSPOSED as fp32 ([3][64][64], wT[m][n*64+k] = W_m[k*64+n]). */
DISTEGNN_API int distegnn_virtual_layer_bwd_simt(int64_t n_nodes, int n_graphs, int A, int C, int Na, unsigned flags,
                                                 const int32_t* batch32, const float* x4, const float* Hn, const float* Xv,
                                                 const float* G, const float* layer_params, const float* wT,
                                                 const float* g_agg_v, const float* g_trans_v, const float* g_vsum,
                                                 float* g_Hn, float* g_xv, float* g_G, float* g_Xv, float* g_layer_params,
                                                 void* stream);

/* Same contract as distegnn_edge_layer_bwd, every tile GEMM as fp32 FMA on the CUDA cores (csrc/edge_layer_bwd.cu): the
 * first backward kernel, kept as the twin of the tensor-core one for cross-checks. */
DISTEGNN_API int distegnn_edge_layer_bwd_simt(int64_t n_nodes, int64_t n_edges, int A, int C, int Na, unsigned flags,
                                              const int32_t* row, const int32_t* col, const float* edge_attr_sorted,
                                              const float* x4, const float* P, const float* Q, const float* layer_params,
                                              const float* g_agg_m, const float* g_agg_x, float* g_P, float* g_Q,
                                              float* g_x4, float* g_layer_params, void* stream);

/* Same contract as distegnn_edge_layer_fwd: the thread-per-row tcgen05 kernel (16 warps per SM, 128 registers per
 * thread; csrc/edge_layer_tc16.cu).  The production symbol runs the column-split flavour (two threads per row, 32
 * warps per SM; csrc/edge_layer_cs.cu); this twin is kept for cross-checks and A/B timing. */
DISTEGNN_API int distegnn_edge_layer_fwd_t16(int64_t n_nodes, int64_t n_edges, int A, int C, int Na, unsigned flags,
                                             const int32_t* row, const int32_t* col, const float* edge_attr_sorted,
                                             const float* x4, const float* P, const float* Q,
                                             const float* layer_params, float* agg_m, float* agg_x, void* stream);

/* Same contract as distegnn_edge_layer_fwd, computed with fp32 FMA on the CUDA cores (no tensor cores).
 * Kept as an independent implementation for cross-checks of the tcgen05 kernel at sizes the CPU oracle
 * cannot reach; not used by FastEGNN.forward. */
DISTEGNN_API int distegnn_edge_layer_fwd_simt(int64_t n_nodes, int64_t n_edges, int A, int C, int Na,
                                              unsigned flags, const int32_t *row, const int32_t *col,
                                              const float *edge_attr_sorted, const float *x4, const float *P,
                                              const float *Q, const float *layer_params, float *agg_m,
                                              float *agg_x, void *stream);

/* Same contract, tensor-core implementation with the 3xTF32 split (earlier production kernel; kept for A/B
 * measurements and as a third independent implementation in the cross-checks). */
DISTEGNN_API int distegnn_edge_layer_fwd_tf32(int64_t n_nodes, int64_t n_edges, int A, int C, int Na,
                                              unsigned flags, const int32_t *row, const int32_t *col,
                                              const float *edge_attr_sorted, const float *x4, const float *P,
                                              const float *Q, const float *layer_params, float *agg_m,
                                              float *agg_x, void *stream);

/* tcgen05 building-block self-test: D[128,64] = A[128,64]·W[64,64]^T on the tensor cores (variant 0:
 * 3xTF32 with A in TMEM, as the fused kernels use it; 2: A in shared memory; 4/6: single-pass TF32). */
DISTEGNN_API int distegnn_selftest_umma(const float *A, const float *W, float *D, int variant, void *stream);

/* 3xTF32 tensor-core twin of distegnn_virtual_layer_fwd (cross-check / A-B timing only). */
DISTEGNN_API int distegnn_virtual_layer_fwd_tf32(int64_t n_nodes, int n_graphs, int A, int C, int Na,
                                                 unsigned flags, const int32_t *batch32, const float *x4,
                                                 const float *Hn, const float *Xv, const float *G,
                                                 const float *layer_params, float *agg_v, float *trans_v,
                                                 float *vsum, void *stream);

/* fp32-FMA twin of distegnn_virtual_layer_fwd (cross-check only). */
DISTEGNN_API int distegnn_virtual_layer_fwd_simt(int64_t n_nodes, int n_graphs, int A, int C, int Na,
                                                 unsigned flags, const int32_t *batch32, const float *x4,
                                                 const float *Hn, const float *Xv, const float *G,
                                                 const float *layer_params, float *agg_v, float *trans_v,
                                                 float *vsum, void *stream);

/* fp32-FMA twin of distegnn_node_layer_fwd (cross-check only). */
DISTEGNN_API int distegnn_node_layer_fwd_simt(int64_t n_nodes, int n_graphs, int A, int C, int Na, unsigned flags,
                                              const int32_t *rowptr, const int32_t *batch32, const float *h,
                                              const float *x4, const float *node_vel, const float *node_attr,
                                              const float *agg_m, const float *agg_x, const float *agg_v,
                                              const float *trans_v, const float *layer_params,
                                              const float *next_layer_params, float *h_out, float *x4_out,
                                              float *P, float *Q, float *Hn, float *node_loc_out, float *vsum,
                                              void *stream);

/* TMA gather4 building-block self-test: out [n_groups][4][box_floats] = rows idx[4g..4g+3] of src [n_rows][64], one
 * cp.async.bulk.tensor ...tile::gather4 per group through a tensor map with box {box_floats, box_rows}; box_floats > 64
 * exercises the zero-filled out-of-bounds tail that gives the rows a padded shared-memory pitch. */
DISTEGNN_API int distegnn_selftest_gather4(const float *src, int64_t n_rows, const int32_t *idx, int n_groups,
                                           int box_floats, int box_rows, float *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DISTEGNN_B200_TESTING_H */
