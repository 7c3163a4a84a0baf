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
    def build_csr(self, edge_index: Tensor, n_nodes: int, validate: bool = True
                  ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        """int64 COO -> int32 CSR by destination.  `validate`: read back the out-of-range counter (one host sync per
        build; builds are cached per edge_index) and raise ValueError like the reference's index assert would;
        validate="defer" hands the device counter back as a fifth value instead (the caller reads it together with the
        data_batch counter after the embed kernel: one pipeline drain for both checks; ids are clamped meanwhile)."""
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
        bad = torch.empty(1, dtype=torch.int32, device=dev) if validate else None
        check(self.lib.distegnn_build_csr(ptr(edge_index), n_nodes, E, ptr(rowptr), ptr(row), ptr(col),
                                          ptr(perm), ptr(ws), ws.numel(), ptr(bad), stream), "build_csr")
        self.launches += (5 if E else 1) + (1 if validate else 0)
        if validate == "defer":
            return rowptr, row, col, perm, (bad if E else None)
        if validate and E and int(bad.item()) != 0:
            raise ValueError(f"edge_index has {int(bad.item())} edge(s) with a node id outside [0, {n_nodes})")
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
              Hn, vsum, n_invalid=None) -> None:
        """`n_invalid`: zeroed int32 [1] device counter of data_batch entries that are unsorted / outside [0,B)."""
        N, B, F, A, Cn, Na = dims
        check(self.lib.distegnn_embed_fwd(N, B, F, A, Cn, Na, ptr(node_feat), ptr(node_loc),
                                          ptr(data_batch), ptr(emb_wt), ptr(emb_b), ptr(layer0), ptr(h),
                                          ptr(x4), ptr(batch32), ptr(P), ptr(Q), ptr(Hn), ptr(vsum),
                                          ptr(n_invalid), self._s(h)), "embed_fwd")
        self.launches += 1 if N else 0

    def edge_layer(self, dims, flags, row, col, ea, x4, P, Q, lp, agg_m, agg_x, n_edges_dev=None) -> None:
        """`n_edges_dev`: int32 [1] on the device with the true edge count when E is only a capacity (CSRGraph built on the
        device without a host round trip)."""
        N, E, A, Cn, Na = dims
        check(self.lib.distegnn_edge_layer_fwd(N, E, A, Cn, Na, flags, ptr(row), ptr(col), ptr(ea),
                                               ptr(x4), ptr(P), ptr(Q), ptr(lp), ptr(agg_m), ptr(agg_x),
                                               ptr(n_edges_dev), self._s(x4)), "edge_layer_fwd")
        self.launches += 1 if E else 0

    def edge_layer_bwd(self, dims, flags, row, col, ea, x4, P, Q, lp, g_agg_m, g_agg_x, g_P, g_Q, g_x4, g_lp,
                       n_edges_dev=None) -> None:
        """Backward of edge_layer: accumulates into g_P, g_Q, g_x4 and the parameter-gradient block g_lp."""
        N, E, A, Cn, Na = dims
        check(self.lib.distegnn_edge_layer_bwd(N, E, A, Cn, Na, flags, ptr(row), ptr(col), ptr(ea), ptr(x4), ptr(P),
                                               ptr(Q), ptr(lp), ptr(g_agg_m), ptr(g_agg_x), ptr(g_P), ptr(g_Q),
                                               ptr(g_x4), ptr(g_lp), ptr(n_edges_dev), self._s(x4)), "edge_layer_bwd")
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

    def virtual_layer(self, dims, flags, batch32, x4, Hn, Xv, G, lp, agg_v, trans_v, vsum) -> None:
        N, B, A, Cn, Na = dims
        check(self.lib.distegnn_virtual_layer_fwd(N, B, A, Cn, Na, flags, ptr(batch32), ptr(x4), ptr(Hn),
                                                  ptr(Xv), ptr(G), ptr(lp), ptr(agg_v), ptr(trans_v),
                                                  ptr(vsum), self._s(x4)), "virtual_layer_fwd")
        self.launches += 1 if N else 0

    def node_layer(self, dims, flags, rowptr, batch32, h, x4, vel, attr, agg_m, agg_x, agg_v, trans_v,
                   lp, lp_next, h_out, x4_out, P, Q, Hn, loc_out, vsum) -> None:
        N, B, A, Cn, Na = dims
        check(self.lib.distegnn_node_layer_fwd(N, B, A, Cn, Na, flags, ptr(rowptr), ptr(batch32), ptr(h),
                                               ptr(x4), ptr(vel), ptr(attr), ptr(agg_m), ptr(agg_x),
                                               ptr(agg_v), ptr(trans_v), ptr(lp), ptr(lp_next),
                                               ptr(h_out), ptr(x4_out), ptr(P), ptr(Q), ptr(Hn),
                                               ptr(loc_out), ptr(vsum), self._s(x4)), "node_layer_fwd")
        self.launches += 1 if N else 0

    def node_layer_bwd(self, dims, flags, rowptr, batch32, h, vel, attr, agg_m, agg_v, lp, lp_next, g_x_out, g_vsum,
                       g_h_out, g_P, g_Q, g_Hn, g_h, g_x, g_agg_x, g_trans_v, g_agg_m, g_agg_v, g_lp, g_lp_next) -> None:
        """Backward of node_layer (csrc/node_layer_bwd.cu): writes g_h, g_x [N,3], g_agg_x, g_trans_v [N,4], g_agg_m, g_agg_v;
        accumulates parameter gradients into g_lp (this layer) and g_lp_next (the projections of the next layer)."""
        N, B, A, Cn, Na = dims
        check(self.lib.distegnn_node_layer_bwd(N, A, Cn, Na, flags, ptr(rowptr), ptr(h), ptr(vel), ptr(attr), ptr(agg_m),
                                               ptr(agg_v), ptr(lp), ptr(lp_next), ptr(g_x_out), ptr(g_vsum), ptr(batch32),
                                               ptr(g_h_out), ptr(g_P), ptr(g_Q), ptr(g_Hn), ptr(g_h), ptr(g_x),
                                               ptr(g_agg_x), ptr(g_trans_v), ptr(g_agg_m), ptr(g_agg_v), ptr(g_lp),
                                               ptr(g_lp_next), self._s(h)), "node_layer_bwd")
        self.launches += 1 if N else 0

    def embed_bwd(self, dims, node_feat, h0, lp0, g_h, g_P, g_Q, g_Hn, g_emb_wt, g_emb_b, g_lp0) -> None:
        """Backward of embed: accumulates g_emb_wt [F,64], g_emb_b [64] and layer 0's projection gradients into g_lp0."""
        N, B, F, A, Cn, Na = dims
        check(self.lib.distegnn_embed_bwd(N, F, A, Cn, Na, ptr(node_feat), ptr(h0), ptr(lp0), ptr(g_h), ptr(g_P), ptr(g_Q),
                                          ptr(g_Hn), ptr(g_emb_wt), ptr(g_emb_b), ptr(g_lp0), self._s(h0)), "embed_bwd")
        self.launches += 1 if N else 0

    def virtual_update(self, dims, flags, vsum, Xv, Hv, lp, lp_next, G, init_loc_mean=None, init_hv0=None,
                       comm: "Optional[Comm]" = None) -> None:
        """Virtual-node update; with `comm` the same kernel first all-reduces vsum over the partitions (NVLink peer
        memory) — the fused form of weighted_average_reduce + update."""
        B, A, Cn, Na = dims
        check(self.lib.distegnn_virtual_update_fwd(B, A, Cn, Na, flags, ptr(vsum), ptr(Xv), ptr(Hv),
                                                   ptr(lp), ptr(lp_next), ptr(G), ptr(init_loc_mean), ptr(init_hv0),
                                                   comm.handle if comm is not None else None, self._s(vsum)),
              "virtual_update_fwd")
        self.launches += 1 if B else 0

    def virtual_update_bwd(self, dims, flags, vsum, Xv, Hv, lp, lp_next, g_Xn, g_Hn, g_G, g_vsum, g_Xv, g_Hv, g_lp,
                           g_lp_next) -> None:
        """Backward of virtual_update (csrc/virtual_update.cu): writes g_vsum, g_Xv, g_Hv; accumulates into g_lp / g_lp_next."""
        B, A, Cn, Na = dims
        check(self.lib.distegnn_virtual_update_bwd(B, A, Cn, Na, flags, ptr(vsum), ptr(Xv), ptr(Hv), ptr(lp), ptr(lp_next),
                                                   ptr(g_Xn), ptr(g_Hn), ptr(g_G), ptr(g_vsum), ptr(g_Xv), ptr(g_Hv),
                                                   ptr(g_lp), ptr(g_lp_next), self._s(vsum)), "virtual_update_bwd")
        self.launches += 1 if B else 0

    def allreduce_packed(self, comm: "Comm", buf: Tensor) -> None:
        """In-place SUM of `buf` over the partitions through the communicator's peer-mapped segments."""
        check(self.lib.distegnn_allreduce_packed(comm.handle, ptr(buf), buf.numel(), self._s(buf)), "allreduce_packed")
        self.launches += 1 if buf.numel() else 0


class Comm:
    """Communicator of the virtual-node sync (C ABI: distegnn_comm_*).  One per (process group, capacity).  The IPC
    handles are all-gathered with torch.distributed (plumbing); the exchange itself is the library's own kernel."""

    def __init__(self, lib, device: torch.device, group, max_slots: int, slot_floats: int):
        import torch.distributed as dist
        self.lib, self.group, self.handle = lib, group, None
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.max_slots, self.slot_floats = int(max_slots), int(slot_floats)
        nb = lib.distegnn_comm_handle_bytes()
        mine = (C.c_ubyte * nb)()
        h = C.c_void_p()
        with torch.cuda.device(device):
            check(lib.distegnn_comm_init(self.rank, self.world, self.max_slots, self.slot_floats, C.byref(h), mine),
                  "comm_init")
            self.handle = h
            gathered = [None] * self.world
            dist.all_gather_object(gathered, bytes(mine), group=group)
            allh = (C.c_ubyte * (nb * self.world)).from_buffer_copy(b"".join(gathered))
            rc = lib.distegnn_comm_connect(self.handle, allh)
            # every rank must agree on the outcome: a half-connected group would dead-lock in the first exchange
            oks = [None] * self.world
            dist.all_gather_object(oks, rc == 0, group=group)
            if not all(oks):
                msg = lib.distegnn_last_error().decode("utf-8", "replace") if rc != 0 else "a peer failed to connect"
                self.destroy()
                raise _lib.DistEGNNError(f"comm_connect failed: {msg}")

    def status(self) -> int:
        v = C.c_int(0)
        check(self.lib.distegnn_comm_status(self.handle, C.byref(v)), "comm_status")
        return int(v.value)

    def destroy(self) -> None:
        """Collective teardown: unmap the peers, wait until every rank has done so, then free the own segment."""
        if self.handle is not None:
            import torch.distributed as dist
            self.lib.distegnn_comm_disconnect(self.handle)
            try:
                dist.barrier(group=self.group)
            except Exception:        # noqa: BLE001 — process group already gone: nothing left to order against
                pass
            self.lib.distegnn_comm_destroy(self.handle)
            self.handle = None


_cuda_backend: Optional[CudaBackend] = None


def cuda_backend() -> CudaBackend:
    global _cuda_backend
    if _cuda_backend is None:
        _cuda_backend = CudaBackend()
    return _cuda_backend
