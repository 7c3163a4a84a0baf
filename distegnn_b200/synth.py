"""Synthetic particle graphs and partitioners for tests and ``bench.py``.

The datasets of the reference are not redistributable/offline, so every workload here is a seeded
synthetic restatement of what the reference's data pipeline hands to ``FastEGNN.forward``:

* point clouds sized like BASELINE.json's configs,
* ``radius_graph(pos, r, loop=False, max_num_neighbors=N)`` as used at
  ``datasets/distribute_graphs.py:43`` (all ordered pairs with ‖Δx‖ < r, both directions),
* ``edge_attr`` = the edge length duplicated into two columns (``distribute_graphs.py:44``),
* ``loc_mean`` = centroid of the *whole* graph, shared by all partitions (``distribute_graphs.py:32``),
* ``split_mode=random`` (``distribute_graphs.py:26-30``) and ``split_mode=kmeans``
  (``distribute_graphs.py:118-143,188-198``) node partitioning with per-partition radius graphs
  (cross-partition edges are dropped, as in the reference).

CPU/numpy only — this is input preparation, not part of the measured path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch


@dataclass
class Workload:
    """Model dims + graph recipe of one BASELINE.json config."""
    name: str
    n_nodes: int
    radius: Optional[float]      # None = fully connected
    degree: float                # expected degree used to size the box
    node_feat_nf: int
    node_attr_nf: int
    edge_attr_nf: int
    virtual_channels: int
    normalize: bool


# BASELINE.json configs
WORKLOADS: Dict[str, Workload] = {
    "nbody100": Workload("nbody100", 100, None, 99.0, 2, 0, 2, 3, True),
    "water3d_10k": Workload("water3d_10k", 10_000, 0.035, 12.2, 2, 0, 2, 3, False),
    "fluid113k": Workload("fluid113k", 113_140, 0.075, 15.1, 3, 2, 2, 5, False),
    "synth1m": Workload("synth1m", 1_000_000, 0.075, 21.0, 3, 2, 2, 8, False),
}


def box_side(n: int, r: float, degree: float) -> float:
    """Side L of the cube such that uniform points have the expected degree: n·(4/3)πr³/L³ = d."""
    return (n * (4.0 / 3.0) * math.pi * r ** 3 / degree) ** (1.0 / 3.0)


def radius_graph_np(pos: np.ndarray, r: float) -> np.ndarray:
    """All ordered pairs (i,j), i≠j, ‖x_i−x_j‖<r → int64 [2,E].  Ordered like PyG's radius_graph
    output is *not* required by FastEGNN (it scatters by edge_index[0]); we emit pairs grouped by
    the second row (the 'col'/source), mimicking radius_graph's sort-by-target convention so that
    edge_index[0] is unsorted — the CSR build must not assume sortedness."""
    from scipy.spatial import cKDTree
    if pos.shape[0] == 0:
        return np.zeros((2, 0), dtype=np.int64)
    tree = cKDTree(pos)
    pairs = tree.query_pairs(r, output_type="ndarray")          # i<j, unique
    if pairs.size == 0:
        return np.zeros((2, 0), dtype=np.int64)
    src = np.concatenate([pairs[:, 0], pairs[:, 1]])
    dst = np.concatenate([pairs[:, 1], pairs[:, 0]])
    order = np.argsort(dst, kind="stable")                       # group by edge_index[1]
    return np.stack([src[order], dst[order]]).astype(np.int64)


def fully_connected_np(n: int) -> np.ndarray:
    """[(i,j) for i for j if i≠j] as the N-body pipeline builds it (process_dataset.py:98-99)."""
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    m = i != j
    return np.stack([i[m], j[m]]).astype(np.int64)


def make_points(w: Workload, seed: int = 0, n_nodes: Optional[int] = None) -> Dict[str, np.ndarray]:
    """Seeded node arrays of one whole (un-partitioned) graph."""
    n = w.n_nodes if n_nodes is None else n_nodes
    rng = np.random.default_rng(seed)
    if w.radius is None:                                         # N-body-like
        pos = rng.normal(0.0, 2.8, size=(n, 3))
        vel = rng.normal(size=(n, 3))
        vel *= 0.5 / np.linalg.norm(vel, axis=1, keepdims=True)
    else:
        side = box_side(n, w.radius, w.degree)
        pos = rng.uniform(0.0, side, size=(n, 3))
        vel = rng.normal(0.0, 0.01, size=(n, 3))
    feat = rng.normal(size=(n, w.node_feat_nf))
    attr = rng.normal(size=(n, w.node_attr_nf))
    return dict(pos=pos.astype(np.float32), vel=vel.astype(np.float32),
                feat=feat.astype(np.float32), attr=attr.astype(np.float32))


def _graph_inputs(pos, vel, feat, attr, loc_mean, radius, edge_attr_nf) -> Dict[str, torch.Tensor]:
    ei = fully_connected_np(pos.shape[0]) if radius is None else radius_graph_np(pos, radius)
    d = np.sqrt(((pos[ei[0]] - pos[ei[1]]) ** 2).sum(-1, dtype=np.float32)).astype(np.float32)
    ea = np.repeat(d[:, None], edge_attr_nf, axis=1)
    n = pos.shape[0]
    return dict(
        node_feat=torch.from_numpy(feat), node_loc=torch.from_numpy(pos),
        node_vel=torch.from_numpy(vel), loc_mean=torch.from_numpy(loc_mean),
        edge_index=torch.from_numpy(ei), data_batch=torch.zeros(n, dtype=torch.long),
        edge_attr=torch.from_numpy(np.ascontiguousarray(ea)),
        node_attr=torch.from_numpy(attr) if attr.shape[1] > 0 else None)


def random_partition(n: int, world_size: int, seed: int = 0) -> List[np.ndarray]:
    """distribute_graphs.py:26-30 — randperm, P−1 chunks of ⌊N/P⌋, remainder to the last."""
    g = torch.Generator().manual_seed(seed)
    idx = torch.randperm(n, generator=g).numpy()
    sizes = [n // world_size] * (world_size - 1)
    sizes.append(n - sum(sizes))
    out, o = [], 0
    for s in sizes:
        out.append(idx[o:o + s])
        o += s
    return out


def kmeans_partition(pos: np.ndarray, world_size: int) -> List[np.ndarray]:
    """distribute_graphs.py:188-198 — sklearn KMeans(n_clusters=P, random_state=0, n_init='auto')
    on float32 positions; partition i = nodes with label i, in index order (``pos[cluster == i]``)."""
    from sklearn.cluster import KMeans
    labels = KMeans(n_clusters=world_size, random_state=0, n_init="auto").fit_predict(
        pos.astype(np.float32))
    return [np.nonzero(labels == i)[0] for i in range(world_size)]


def make_partitions(w: Workload, world_size: int = 1, split_mode: str = "random", seed: int = 0,
                    n_nodes: Optional[int] = None, only_rank: Optional[int] = None
                    ) -> List[Optional[Dict[str, torch.Tensor]]]:
    """One input dict per partition (= per rank), each with the forward() argument names.

    ``only_rank`` builds the (expensive) radius graph for that rank only and leaves ``None``
    elsewhere — every rank of a torchrun job calls this with its own rank and the same seed.
    """
    pts = make_points(w, seed, n_nodes)
    n = pts["pos"].shape[0]
    loc_mean = pts["pos"].mean(axis=0, keepdims=True, dtype=np.float64).astype(np.float32)
    if world_size == 1:
        chunks = [np.arange(n)]
    elif split_mode == "random":
        chunks = random_partition(n, world_size, seed)
    elif split_mode == "kmeans":
        chunks = kmeans_partition(pts["pos"], world_size)
    else:
        raise ValueError(f"unsupported split_mode {split_mode!r} (random|kmeans)")
    out: List[Optional[Dict[str, torch.Tensor]]] = []
    for r, idx in enumerate(chunks):
        if only_rank is not None and r != only_rank:
            out.append(None)
            continue
        out.append(_graph_inputs(pts["pos"][idx], pts["vel"][idx], pts["feat"][idx], pts["attr"][idx],
                                 loc_mean, w.radius, w.edge_attr_nf))
    return out
