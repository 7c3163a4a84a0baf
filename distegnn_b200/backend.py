"""The one place where torch tensors become raw device pointers for the C ABI.

`CudaBackend` is the product path.  It has no CPU branch: tensors must live on a CUDA device and the
shared library must load.  (tests/ contains a pure-torch stand-in with the same method names that is
used ONLY to exercise the host-side sequencing and the multi-partition all-reduce logic under gloo on
machines without a GPU; it is never importable from the package.)
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check, ptr

Tensor = torch.Tensor


class CudaBackend:
    name = "cuda-sm100a"

    def __init__(self) -> None:
        self.lib = _lib.load()
        self.launches = 0          # kernels of ours enqueued (bench.py reports it as gpu_launches)

    # ---- helpers -------------------------------------------------------------------------------
    @staticmethod
    def _s(t: Tensor) -> int:
        if not t.is_cuda:
            raise _lib.DistEGNNError("distegnn_b200 has no CPU path: tensors must be on a CUDA device")
        return _lib.stream_ptr(t.device)

    # ---- graph preprocessing -------------------------------------------------------------------
    def build_csr(self, edge_index: Tensor, n_nodes: int) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        E = int(edge_index.shape[1])
        dev = edge_index.device
        stream = self._s(edge_index)
        rowptr = torch.empty(n_nodes + 1, dtype=torch.int32, device=dev)
        row = torch.empty(E, dtype=torch.int32, device=dev)
        col = torch.empty(E, dtype=torch.int32, device=dev)
        perm = torch.empty(E, dtype=torch.int32, device=dev)
        nbytes = C.c_int64(0)
        check(self.lib.distegnn_csr_workspace_bytes(n_nodes, E, C.byref(nbytes)), "csr_workspace_bytes")
        ws = torch.empty(max(int(nbytes.value), 1), dtype=torch.uint8, device=dev)
        check(self.lib.distegnn_build_csr(ptr(edge_index), n_nodes, E, ptr(rowptr), ptr(row), ptr(col),
                                          ptr(perm), ptr(ws), ws.numel(), stream), "build_csr")
        self.launches += 5 if E else 1
        return rowptr, row, col, perm

    def gather_rows(self, src: Tensor, perm: Tensor) -> Tensor:
        dst = torch.empty_like(src)
        if src.numel():
            check(self.lib.distegnn_gather_rows(ptr(src), ptr(perm), src.shape[0], src.shape[1], ptr(dst),
                                                self._s(src)), "gather_rows")
            self.launches += 1
        return dst

    # ---- layer stages --------------------------------------------------------------------------
    def embed(self, dims, node_feat, node_loc, data_batch, emb_wt, emb_b, layer0, h, x4, batch32, P, Q,
              Hn, vsum) -> None:
        N, B, F, A, Cn, Na = dims
        check(self.lib.distegnn_embed_fwd(N, B, F, A, Cn, Na, ptr(node_feat), ptr(node_loc),
                                          ptr(data_batch), ptr(emb_wt), ptr(emb_b), ptr(layer0), ptr(h),
                                          ptr(x4), ptr(batch32), ptr(P), ptr(Q), ptr(Hn), ptr(vsum),
                                          self._s(h)), "embed_fwd")
        self.launches += 1 if N else 0

    def embed_simt(self, dims, node_feat, node_loc, data_batch, emb_wt, emb_b, layer0, h, x4, batch32, P, Q,
                   Hn, vsum) -> None:
        """fp32-FMA twin of embed (cross-check only)."""
        N, B, F, A, Cn, Na = dims
        check(self.lib.distegnn_embed_fwd_simt(N, B, F, A, Cn, Na, ptr(node_feat), ptr(node_loc),
                                               ptr(data_batch), ptr(emb_wt), ptr(emb_b), ptr(layer0), ptr(h),
                                               ptr(x4), ptr(batch32), ptr(P), ptr(Q), ptr(Hn), ptr(vsum),
                                               self._s(h)), "embed_fwd_simt")

    def edge_layer(self, dims, flags, row, col, ea, x4, P, Q, lp, agg_m, agg_x) -> None:
        N, E, A, Cn, Na = dims
        check(self.lib.distegnn_edge_layer_fwd(N, E, A, Cn, Na, flags, ptr(row), ptr(col), ptr(ea),
                                               ptr(x4), ptr(P), ptr(Q), ptr(lp), ptr(agg_m), ptr(agg_x),
                                               self._s(x4)), "edge_layer_fwd")
        self.launches += 1 if E else 0

    def edge_layer_bwd(self, dims, flags, row, col, ea, x4, P, Q, lp, g_agg_m, g_agg_x, g_P, g_Q, g_x4, g_lp) -> None:
        """Backward of edge_layer: accumulates into g_P, g_Q, g_x4 and the parameter-gradient block g_lp."""
        N, E, A, Cn, Na = dims
        check(self.lib.distegnn_edge_layer_bwd(N, E, A, Cn, Na, flags, ptr(row), ptr(col), ptr(ea), ptr(x4), ptr(P),
                                               ptr(Q), ptr(lp), ptr(g_agg_m), ptr(g_agg_x), ptr(g_P), ptr(g_Q),
                                               ptr(g_x4), ptr(g_lp), self._s(x4)), "edge_layer_bwd")
        self.launches += 1 if E else 0

    def virtual_bwd_prepare(self, A, Cn, Na, lp) -> "torch.Tensor":
        """fp16 hi/lo operand images of the virtual stage's weights and their transposes (96 KB) for virtual_layer_bwd."""
        import torch
        img = torch.empty(6 * 2 * 64 * 64, dtype=torch.float16, device=lp.device)
        check(self.lib.distegnn_virtual_bwd_prepare(A, Cn, Na, ptr(lp), ptr(img), self._s(lp)), "virtual_bwd_prepare")
        self.launches += 1
        return img

    def virtual_layer_bwd(self, dims, flags, batch32, x4, Hn, Xv, G, lp, wimg, g_agg_v, g_trans_v, g_vsum, g_Hn, g_xv,
                          g_G, g_Xv, g_lp) -> None:
        """Backward of virtual_layer (tcgen05): writes g_Hn, g_xv; accumulates into g_G, g_Xv and the parameter gradients.
        `wimg` comes from virtual_bwd_prepare(lp)."""
        N, B, A, Cn, Na = dims
        check(self.lib.distegnn_virtual_layer_bwd(N, B, A, Cn, Na, flags, ptr(batch32), ptr(x4), ptr(Hn), ptr(Xv),
                                                  ptr(G), ptr(lp), ptr(wimg), ptr(g_agg_v), ptr(g_trans_v), ptr(g_vsum),
                                                  ptr(g_Hn), ptr(g_xv), ptr(g_G), ptr(g_Xv), ptr(g_lp), self._s(x4)),
              "virtual_layer_bwd")
        self.launches += 1 if N else 0

    def virtual_layer_bwd_simt(self, dims, flags, batch32, x4, Hn, Xv, G, lp, wT, g_agg_v, g_trans_v, g_vsum, g_Hn, g_xv,
                               g_G, g_Xv, g_lp) -> None:
        """fp32-FMA twin of virtual_layer_bwd (cross-check only); wT = the three matrices transposed, fp32."""
        N, B, A, Cn, Na = dims
        check(self.lib.distegnn_virtual_layer_bwd_simt(N, B, A, Cn, Na, flags, ptr(batch32), ptr(x4), ptr(Hn), ptr(Xv),
                                                       ptr(G), ptr(lp), ptr(wT), ptr(g_agg_v), ptr(g_trans_v), ptr(g_vsum),
                                                       ptr(g_Hn), ptr(g_xv), ptr(g_G), ptr(g_Xv), ptr(g_lp), self._s(x4)),
              "virtual_layer_bwd_simt")

    @staticmethod
    def _grid_host(grid):
        import ctypes as C
        origin, cell, dims = grid
        return (C.c_float * 3)(*origin), float(cell), (C.c_int32 * 3)(*dims)

    def radius_count(self, N, x4, batch32, order32, cell_start, grid, r, loop, deg) -> None:
        """Neighbour counts of the on-device radius graph (csrc/radius_graph.cu)."""
        import ctypes as C
        o, cell, d = self._grid_host(grid)
        check(self.lib.distegnn_radius_count(N, ptr(x4), ptr(batch32), ptr(order32), ptr(cell_start),
                                             C.cast(o, C.c_void_p), cell, C.cast(d, C.c_void_p), float(r), int(loop),
                                             ptr(deg), self._s(x4)), "radius_count")
        self.launches += 1 if N else 0

    def radius_fill(self, N, x4, batch32, order32, cell_start, grid, r, loop, rowptr, row, col, dist) -> None:
        import ctypes as C
        o, cell, d = self._grid_host(grid)
        check(self.lib.distegnn_radius_fill(N, ptr(x4), ptr(batch32), ptr(order32), ptr(cell_start),
                                            C.cast(o, C.c_void_p), cell, C.cast(d, C.c_void_p), float(r), int(loop),
                                            ptr(rowptr), ptr(row), ptr(col), ptr(dist), self._s(x4)), "radius_fill")
        self.launches += 1 if N else 0

    def edge_layer_bwd_simt(self, dims, flags, row, col, ea, x4, P, Q, lp, g_agg_m, g_agg_x, g_P, g_Q, g_x4, g_lp) -> None:
        """fp32-FMA twin of edge_layer_bwd (cross-check only)."""
        N, E, A, Cn, Na = dims
        check(self.lib.distegnn_edge_layer_bwd_simt(N, E, A, Cn, Na, flags, ptr(row), ptr(col), ptr(ea), ptr(x4),
                                                    ptr(P), ptr(Q), ptr(lp), ptr(g_agg_m), ptr(g_agg_x), ptr(g_P),
                                                    ptr(g_Q), ptr(g_x4), ptr(g_lp), self._s(x4)), "edge_layer_bwd_simt")

    def edge_layer_t16(self, dims, flags, row, col, ea, x4, P, Q, lp, agg_m, agg_x) -> None:
        """thread-per-row tcgen05 twin of edge_layer (cross-check / A-B timing only)."""
        N, E, A, Cn, Na = dims
        check(self.lib.distegnn_edge_layer_fwd_t16(N, E, A, Cn, Na, flags, ptr(row), ptr(col), ptr(ea),
                                                   ptr(x4), ptr(P), ptr(Q), ptr(lp), ptr(agg_m),
                                                   ptr(agg_x), self._s(x4)), "edge_layer_fwd_t16")

    def edge_layer_simt(self, dims, flags, row, col, ea, x4, P, Q, lp, agg_m, agg_x) -> None:
        """fp32-FMA twin of edge_layer (cross-check only; FastEGNN.forward never calls it)."""
        N, E, A, Cn, Na = dims
        check(self.lib.distegnn_edge_layer_fwd_simt(N, E, A, Cn, Na, flags, ptr(row), ptr(col), ptr(ea),
                                                    ptr(x4), ptr(P), ptr(Q), ptr(lp), ptr(agg_m),
                                                    ptr(agg_x), self._s(x4)), "edge_layer_fwd_simt")

    def virtual_layer(self, dims, flags, batch32, x4, Hn, Xv, G, lp, agg_v, trans_v, vsum) -> None:
        N, B, A, Cn, Na = dims
        check(self.lib.distegnn_virtual_layer_fwd(N, B, A, Cn, Na, flags, ptr(batch32), ptr(x4), ptr(Hn),
                                                  ptr(Xv), ptr(G), ptr(lp), ptr(agg_v), ptr(trans_v),
                                                  ptr(vsum), self._s(x4)), "virtual_layer_fwd")
        self.launches += 1 if N else 0

    def edge_layer_tf32(self, dims, flags, row, col, ea, x4, P, Q, lp, agg_m, agg_x) -> None:
        """3xTF32 tensor-core twin of edge_layer (cross-check / A-B timing only)."""
        N, E, A, Cn, Na = dims
        check(self.lib.distegnn_edge_layer_fwd_tf32(N, E, A, Cn, Na, flags, ptr(row), ptr(col), ptr(ea),
                                                    ptr(x4), ptr(P), ptr(Q), ptr(lp), ptr(agg_m),
                                                    ptr(agg_x), self._s(x4)), "edge_layer_fwd_tf32")

    def virtual_layer_cs(self, dims, flags, batch32, x4, Hn, Xv, G, lp, agg_v, trans_v, vsum) -> None:
        """thread-per-row tcgen05 twin of virtual_layer (cross-check / A-B timing only)."""
        N, B, A, Cn, Na = dims
        check(self.lib.distegnn_virtual_layer_fwd_cs(N, B, A, Cn, Na, flags, ptr(batch32), ptr(x4), ptr(Hn),
                                                      ptr(Xv), ptr(G), ptr(lp), ptr(agg_v), ptr(trans_v),
                                                      ptr(vsum), self._s(x4)), "virtual_layer_fwd_cs")

    def virtual_layer_tf32(self, dims, flags, batch32, x4, Hn, Xv, G, lp, agg_v, trans_v, vsum) -> None:
        """3xTF32 tensor-core twin of virtual_layer (cross-check / A-B timing only)."""
        N, B, A, Cn, Na = dims
        check(self.lib.distegnn_virtual_layer_fwd_tf32(N, B, A, Cn, Na, flags, ptr(batch32), ptr(x4), ptr(Hn),
                                                       ptr(Xv), ptr(G), ptr(lp), ptr(agg_v), ptr(trans_v),
                                                       ptr(vsum), self._s(x4)), "virtual_layer_fwd_tf32")

    def virtual_layer_simt(self, dims, flags, batch32, x4, Hn, Xv, G, lp, agg_v, trans_v, vsum) -> None:
        """fp32-FMA twin of virtual_layer (cross-check only)."""
        N, B, A, Cn, Na = dims
        check(self.lib.distegnn_virtual_layer_fwd_simt(N, B, A, Cn, Na, flags, ptr(batch32), ptr(x4), ptr(Hn),
                                                       ptr(Xv), ptr(G), ptr(lp), ptr(agg_v), ptr(trans_v),
                                                       ptr(vsum), self._s(x4)), "virtual_layer_fwd_simt")

    def node_layer(self, dims, flags, rowptr, batch32, h, x4, vel, attr, agg_m, agg_x, agg_v, trans_v,
                   lp, lp_next, h_out, x4_out, P, Q, Hn, loc_out, vsum) -> None:
        N, B, A, Cn, Na = dims
        check(self.lib.distegnn_node_layer_fwd(N, B, A, Cn, Na, flags, ptr(rowptr), ptr(batch32), ptr(h),
                                               ptr(x4), ptr(vel), ptr(attr), ptr(agg_m), ptr(agg_x),
                                               ptr(agg_v), ptr(trans_v), ptr(lp), ptr(lp_next),
                                               ptr(h_out), ptr(x4_out), ptr(P), ptr(Q), ptr(Hn),
                                               ptr(loc_out), ptr(vsum), self._s(x4)), "node_layer_fwd")
        self.launches += 1 if N else 0

    def node_layer_simt(self, dims, flags, rowptr, batch32, h, x4, vel, attr, agg_m, agg_x, agg_v, trans_v,
                        lp, lp_next, h_out, x4_out, P, Q, Hn, loc_out, vsum) -> None:
        """fp32-FMA twin of node_layer (cross-check only)."""
        N, B, A, Cn, Na = dims
        check(self.lib.distegnn_node_layer_fwd_simt(N, B, A, Cn, Na, flags, ptr(rowptr), ptr(batch32), ptr(h),
                                                    ptr(x4), ptr(vel), ptr(attr), ptr(agg_m), ptr(agg_x),
                                                    ptr(agg_v), ptr(trans_v), ptr(lp), ptr(lp_next),
                                                    ptr(h_out), ptr(x4_out), ptr(P), ptr(Q), ptr(Hn),
                                                    ptr(loc_out), ptr(vsum), self._s(x4)), "node_layer_fwd_simt")

    def virtual_update(self, dims, flags, vsum, Xv, Hv, lp, lp_next, G) -> None:
        B, A, Cn, Na = dims
        check(self.lib.distegnn_virtual_update_fwd(B, A, Cn, Na, flags, ptr(vsum), ptr(Xv), ptr(Hv),
                                                   ptr(lp), ptr(lp_next), ptr(G), self._s(vsum)),
              "virtual_update_fwd")
        self.launches += 1 if B else 0


_cuda_backend: Optional[CudaBackend] = None


def cuda_backend() -> CudaBackend:
    global _cuda_backend
    if _cuda_backend is None:
        _cuda_backend = CudaBackend()
    return _cuda_backend
