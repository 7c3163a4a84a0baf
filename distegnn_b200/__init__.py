"""distegnn_b200 — B200-native (sm_100a) implementation of the DistEGNN hot path.

Public surface mirrors the reference: ``from distegnn_b200 import FastEGNN`` is a drop-in for
``from models.FastEGNN import FastEGNN``.
"""
from .fast_egnn import E_GCL_vel, FastEGNN  # noqa: F401
from .graph import radius_graph, split_large_graph_random  # noqa: F401  (on-device graph construction)
from .loss import train_loss  # noqa: F401  (fused weighted-MSE + MMD loss of the training step)
from .partition import kmeans_labels, radius_graph_csr, split_large_graph  # noqa: F401  (CSR out, no host round trip)

__all__ = ["FastEGNN", "E_GCL_vel", "radius_graph", "radius_graph_csr", "kmeans_labels", "split_large_graph",
           "split_large_graph_random", "train_loss"]
