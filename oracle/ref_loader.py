"""Import the reference's own ``FastEGNN`` from ``oracle/_ref`` (installed by ``oracle/build_ref.py``).

Test/bench infrastructure — never imported by the product package.  The one thing added is a stand-in for
``torch_geometric.nn.global_mean_pool`` (PyG 2.6.1 is pinned by the reference's requirements.txt:15 and is not
in this image): scatter-sum by graph id divided by ``count.clamp(min=1)``, the published semantics of
``global_mean_pool`` (call sites FastEGNN.py:193, 222, 258).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_FILE = os.path.join(HERE, "_ref", "models", "FastEGNN.py")


def global_mean_pool(x, batch, size=None):
    n = int(batch.max().item()) + 1 if size is None else size
    tot = x.new_zeros((n, x.size(1)))
    tot.index_add_(0, batch, x)
    cnt = torch.bincount(batch, minlength=n).clamp(min=1).to(x.dtype)
    return tot / cnt.unsqueeze(-1)


def available() -> bool:
    return os.path.exists(REF_FILE)


_cls = None


def load_reference():
    """The reference ``FastEGNN`` class (unmodified source), or raises FileNotFoundError."""
    global _cls
    if _cls is not None:
        return _cls
    if not available():
        raise FileNotFoundError(f"{REF_FILE} not found — run `python oracle/build_ref.py` where /root/reference exists")
    if "torch_geometric" not in sys.modules:
        tg = types.ModuleType("torch_geometric")
        tgnn = types.ModuleType("torch_geometric.nn")
        tgnn.global_mean_pool = global_mean_pool
        tg.nn = tgnn
        sys.modules["torch_geometric"] = tg
        sys.modules["torch_geometric.nn"] = tgnn
    spec = importlib.util.spec_from_file_location("_distegnn_reference_fastegnn", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cls = mod.FastEGNN
    return _cls


def reference_forward(sd, node_feat, node_loc, node_vel, loc_mean, edge_index, data_batch, edge_attr, node_attr,
                      normalize: bool = False, n_layers: int = 4):
    """One forward of the reference module (world_size=1) holding the weights `sd` (reference state_dict keys)."""
    cls = load_reference()
    F = sd["embedding_in.weight"].shape[1]
    Hh = sd["embedding_in.weight"].shape[0]
    Cn = sd["virtual_node_feat"].shape[2]
    A = sd["gcl_0.edge_mlp.0.weight"].shape[1] - 2 * Hh - 1
    Na = sd["gcl_0.node_mlp.0.weight"].shape[1] - 3 * Hh
    m = cls(node_feat_nf=F, node_attr_nf=Na, edge_attr_nf=A, hidden_nf=Hh, virtual_channels=Cn, world_size=1,
            n_layers=n_layers, normalize=normalize)
    m.load_state_dict(sd)
    m = m.to(node_loc.dtype).eval()
    with torch.no_grad():
        return m(node_feat, node_loc, node_vel, loc_mean, edge_index, data_batch, edge_attr,
                 node_attr if Na > 0 else None)
