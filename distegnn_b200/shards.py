"""Input wire format: a binary, memory-mappable shard per (graph, rank) that already holds the graph in
the form the kernels consume — CSR by destination row with int32 ids and the edge attributes in CSR order — instead of
the reference's pickled list of PyG ``Data(x, pos, vel, attr, target, loc_mean, edge_index, edge_attr, special_nodes)``
(`datasets/process_dataset.py:114-115`, `datasets/distribute_graphs.py:46-49`) that `DataLoader` collates on the host.

A shard is ONE file: a JSON header (array name -> dtype, shape, byte offset; 64-byte aligned payloads) followed by the raw
little-endian arrays.  `read_shard` memory-maps it; `Shard.pinned()` gives page-locked host tensors for asynchronous H2D;
`Shard.to(device)` returns the keyword arguments of `FastEGNN.forward`, with `edge_index` replaced by a `CSRGraph` — the
model then skips the COO→CSR radix sort and the edge-attribute permutation, and the graph crosses PCIe as 4 bytes per edge
(`col`) plus `rowptr` instead of 16 bytes per edge (`int64 [2,E]`).
"""
from __future__ import annotations

import json
import struct
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

Tensor = torch.Tensor
MAGIC = b"DEGNNSH1"
_ALIGN = 64


@dataclass
class CSRGraph:
    """A graph already sorted by destination row: `rowptr` int32 [N+1], `col` int32 [E]; `row` (int32 [E], the expanded
    destination ids the edge kernels read) is derived on first use.  Accepted by `FastEGNN.forward` in place of the
    int64 `edge_index`; `edge_attr` passed next to it must already be in this edge order."""
    rowptr: Tensor
    col: Tensor
    row: Optional[Tensor] = None
    # graphs built on the device without a host round trip (partition.radius_graph_csr(capacity=...)): `col` / `row` /
    # edge_attr are CAPACITY-sized, the true edge count lives on the device and the kernels read it there
    n_edges_dev: Optional[Tensor] = None
    info: Optional[Tensor] = None

    def __post_init__(self):
        self._checked = None
        for name, t in (("rowptr", self.rowptr), ("col", self.col), ("row", self.row)):
            if t is None:
                continue
            if t.dtype != torch.int32 or t.dim() != 1:
                raise ValueError(f"CSRGraph.{name} must be a 1-D int32 tensor (got {t.dtype}, shape {tuple(t.shape)}); "
                                 "convert explicitly, e.g. torch.cumsum(...).to(torch.int32)")
            if t.device != self.rowptr.device:
                raise ValueError("CSRGraph tensors must live on one device")
        if self.rowptr.numel() < 1:
            raise ValueError("CSRGraph.rowptr must have N+1 >= 1 entries")
        if self.row is not None and self.row.shape != self.col.shape:
            raise ValueError("CSRGraph.row and .col must have the same length")

    def validate(self, device=None) -> None:
        """Content checks the kernels rely on (monotone rowptr ending at E, column ids in range); one host sync, done
        once per CSRGraph object."""
        if device is not None and self.rowptr.device != device:
            raise ValueError(f"CSRGraph lives on {self.rowptr.device}, the node tensors on {device}")
        if self._checked or self.n_edges_dev is not None:     # device-built: valid by construction, count not on the host
            return
        N, E = self.num_nodes, self.num_edges
        ok = int(self.rowptr[0]) == 0 and int(self.rowptr[-1]) == E
        if ok and N > 0:
            ok = bool((self.rowptr[1:] >= self.rowptr[:-1]).all())
        if ok and E > 0:
            ok = N > 0 and int(self.col.min()) >= 0 and int(self.col.max()) < N
        if not ok:
            raise ValueError("CSRGraph is not a valid CSR: rowptr must be non-decreasing from 0 to E and col in [0, N)")
        self._checked = True

    def overflowed(self) -> bool:
        """Device-built graph only: did the edge count exceed the capacity (then edges are missing)?  Synchronises."""
        return self.info is not None and int(self.info[1].item()) != 0

    @property
    def num_nodes(self) -> int:
        return int(self.rowptr.shape[0]) - 1

    @property
    def num_edges(self) -> int:
        return int(self.col.shape[0])

    def rows(self) -> Tensor:
        if self.row is None:
            deg = (self.rowptr[1:] - self.rowptr[:-1]).to(torch.int64)
            self.row = torch.repeat_interleave(torch.arange(self.num_nodes, device=self.col.device, dtype=torch.int32), deg,
                                               output_size=self.num_edges)
        return self.row

    def edge_index(self) -> Tensor:
        """int64 [2,E] in the reference's convention (edge_index[0] = destination)."""
        return torch.stack([self.rows().to(torch.int64), self.col.to(torch.int64)])

    @staticmethod
    def from_edge_index(edge_index: Tensor, n_nodes: int, edge_attr: Optional[Tensor] = None):
        """Host-side construction (stable sort by destination) -> (CSRGraph, edge_attr in CSR order)."""
        row64 = edge_index[0]
        perm = torch.argsort(row64, stable=True)
        rowptr = torch.zeros(n_nodes + 1, dtype=torch.int32, device=edge_index.device)
        rowptr[1:] = torch.cumsum(torch.bincount(row64, minlength=n_nodes), 0).to(torch.int32)
        g = CSRGraph(rowptr, edge_index[1][perm].to(torch.int32).contiguous(), row64[perm].to(torch.int32).contiguous())
        return g, (None if edge_attr is None else edge_attr[perm].contiguous())


def write_shard(path: str, arrays: Dict[str, np.ndarray]) -> None:
    """arrays: name -> numpy array (any of float32 / int32 / int64 / bool); written in the given order."""
    meta, off = {}, 0
    for k, v in arrays.items():
        v = np.ascontiguousarray(v)
        off = (off + _ALIGN - 1) // _ALIGN * _ALIGN
        meta[k] = dict(dtype=str(v.dtype), shape=list(v.shape), offset=off)
        off += v.nbytes
    header = json.dumps(meta).encode()
    base = (len(MAGIC) + 8 + len(header) + _ALIGN - 1) // _ALIGN * _ALIGN
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<II", len(header), base))
        f.write(header)
        for k, v in arrays.items():
            f.seek(base + meta[k]["offset"])
            f.write(np.ascontiguousarray(v).tobytes())


def shard_from_forward_inputs(inp: Dict[str, Optional[Tensor]], target: Optional[Tensor] = None) -> Dict[str, np.ndarray]:
    """`FastEGNN.forward` keyword arguments (host tensors, int64 edge_index) -> shard arrays (CSR built here, once)."""
    n = int(inp["node_loc"].shape[0])
    g, ea = CSRGraph.from_edge_index(inp["edge_index"], n, inp.get("edge_attr"))
    out = dict(node_feat=inp["node_feat"].numpy(), node_loc=inp["node_loc"].numpy(), node_vel=inp["node_vel"].numpy(),
               loc_mean=inp["loc_mean"].numpy(), data_batch=inp["data_batch"].numpy().astype(np.int32),
               rowptr=g.rowptr.numpy(), col=g.col.numpy())
    if ea is not None:
        out["edge_attr"] = ea.numpy()
    if inp.get("node_attr") is not None:
        out["node_attr"] = inp["node_attr"].numpy()
    if target is not None:
        out["target"] = target.numpy()
    return out


class Shard:
    def __init__(self, tensors: Dict[str, Tensor]):
        self.t = tensors

    def pinned(self) -> "Shard":
        return Shard({k: (v if v.is_pinned() else v.clone().pin_memory()) for k, v in self.t.items()})

    def nbytes(self) -> int:
        return sum(v.numel() * v.element_size() for v in self.t.values())

    def to(self, device, non_blocking: bool = True) -> Dict[str, object]:
        """-> keyword arguments for FastEGNN.forward (`edge_index` is a CSRGraph on `device`)."""
        d = {k: v.to(device, non_blocking=non_blocking) for k, v in self.t.items()}
        return dict(node_feat=d["node_feat"], node_loc=d["node_loc"], node_vel=d["node_vel"], loc_mean=d["loc_mean"],
                    edge_index=CSRGraph(d["rowptr"], d["col"]), data_batch=d["data_batch"].to(torch.int64),
                    edge_attr=d.get("edge_attr"), node_attr=d.get("node_attr"))


def read_shard(path: str) -> Shard:
    """Memory-map a shard; the tensors alias the mapping (copy-on-write) until `.pinned()` / `.to()`."""
    with open(path, "rb") as f:
        if f.read(len(MAGIC)) != MAGIC:
            raise ValueError(f"{path}: not a distegnn_b200 shard")
        hlen, base = struct.unpack("<II", f.read(8))
        meta = json.loads(f.read(hlen).decode())
    mm = np.memmap(path, dtype=np.uint8, mode="c")
    out = {}
    for k, m in meta.items():
        n = int(np.prod(m["shape"])) if m["shape"] else 1
        a = np.frombuffer(mm, dtype=np.dtype(m["dtype"]), count=n, offset=base + m["offset"]).reshape(m["shape"])
        out[k] = torch.from_numpy(a)
    return Shard(out)
