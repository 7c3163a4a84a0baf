"""Loader around the shard format on CPU: PyG-style collation on CSR shards, the reference's same-seed
sampling order, and the model fed from the loader vs the int64 edge_index batch."""
import numpy as np
import pytest
import torch

from distegnn_b200 import FastEGNN, synth
from distegnn_b200.loader import ShardLoader, collate
from distegnn_b200.shards import read_shard, shard_from_forward_inputs, write_shard
from oracle import fastegnn_oracle as orc
from tests.shadow_backend import ShadowBackend


def _graphs(k, seed=0):
    w = synth.WORKLOADS["nbody100"]
    out = []
    for i in range(k):
        p = synth.make_partitions(w, n_nodes=9 + 3 * i, seed=seed + i)[0]
        p["target"] = p["node_loc"] + 0.1 * p["node_vel"]
        out.append(p)
    return out


def _write(tmp_path, graphs):
    paths = []
    for i, g in enumerate(graphs):
        p = str(tmp_path / f"g{i}.shard")
        inp = {k: v for k, v in g.items() if k != "target"}
        write_shard(p, shard_from_forward_inputs(inp, target=g["target"]))
        paths.append(p)
    return paths


def _pyg_collate(graphs):
    """What torch_geometric's Batch.from_data_list does with the reference's Data objects (main.py:184-190)."""
    off, ei, batch = 0, [], []
    for i, g in enumerate(graphs):
        ei.append(g["edge_index"] + off)
        batch.append(torch.full((g["node_loc"].shape[0],), i, dtype=torch.long))
        off += g["node_loc"].shape[0]
    cat = lambda k: torch.cat([g[k] for g in graphs], 0)
    return dict(node_feat=cat("node_feat"), node_loc=cat("node_loc"), node_vel=cat("node_vel"), loc_mean=cat("loc_mean"),
                edge_index=torch.cat(ei, 1), data_batch=torch.cat(batch), edge_attr=cat("edge_attr"), node_attr=None), cat("target")


def test_collate_is_pyg_collation_on_csr(tmp_path):
    graphs = _graphs(3)
    paths = _write(tmp_path, graphs)
    host = collate([read_shard(p) for p in paths])
    want, target = _pyg_collate(graphs)
    assert torch.equal(host["node_loc"], want["node_loc"]) and torch.equal(host["target"], target)
    assert torch.equal(host["data_batch"].long(), want["data_batch"]) and torch.equal(host["loc_mean"], want["loc_mean"])
    assert host["ptr"].tolist() == [0, 9, 21, 36]
    # the concatenated CSR is the batch's graph: same (row, col, attr) multiset, rows ascending, nothing re-sorted
    rp, col = host["rowptr"].long(), host["col"].long()
    assert rp[0] == 0 and rp[-1] == col.numel() == want["edge_index"].shape[1] and bool((rp[1:] >= rp[:-1]).all())
    row = torch.repeat_interleave(torch.arange(rp.numel() - 1), rp[1:] - rp[:-1])
    key = lambda r, c, a: sorted(zip(r.tolist(), c.tolist(), [tuple(x) for x in a.tolist()]))
    assert key(row, col, host["edge_attr"]) == key(want["edge_index"][0], want["edge_index"][1], want["edge_attr"])


def test_loader_order_is_the_same_seed_random_sampler(tmp_path):
    from torch.utils.data import RandomSampler
    paths = _write(tmp_path, _graphs(7))
    gen = torch.Generator()
    gen.manual_seed(43)
    want = list(RandomSampler(range(7), replacement=False, generator=gen))       # main.py:185-188
    a = ShardLoader(paths, batch_size=2, shuffle=True, seed=43, drop_last=True)
    b = ShardLoader(paths, batch_size=2, shuffle=True, seed=43, drop_last=True)  # "another rank"
    assert len(a) == 3
    seen = []
    for (ka, ea), (kb, eb) in zip(a, b):
        assert torch.equal(ka["node_loc"], kb["node_loc"])                       # every rank walks the same graphs
        assert ea["n_graphs"] == 2
        seen.append(ka["node_loc"].shape[0])
    sizes = [9 + 3 * i for i in range(7)]
    assert seen == [sizes[want[0]] + sizes[want[1]], sizes[want[2]] + sizes[want[3]], sizes[want[4]] + sizes[want[5]]]


def test_model_from_loader_matches_edge_index_batch(tmp_path):
    graphs = _graphs(3, seed=5)
    paths = _write(tmp_path, graphs)
    sd = orc.init_state_dict(2, 0, 2, 64, 3, 2, seed=1, coord_gain=0.05)
    m = FastEGNN(node_feat_nf=2, node_attr_nf=0, edge_attr_nf=2, hidden_nf=64, virtual_channels=3, world_size=1,
                 n_layers=2, normalize=True)
    m.load_state_dict(sd)
    m._backend = ShadowBackend()
    want_in, target = _pyg_collate(graphs)
    with torch.no_grad():
        want, wantX = m(**want_in)
        (kwargs, extras), = list(ShardLoader(paths, batch_size=3, shuffle=False))
        got, gotX = m(**kwargs)
    assert extras["node_counts"] == [9, 12, 15] and torch.equal(extras["target"], target)
    assert float((got - want).abs().max()) <= 1e-6 and float((gotX - wantX).abs().max()) <= 1e-6
    ref, refX = orc.forward(sd, **want_in, normalize=True)
    assert float((got - ref).abs().max()) <= 1e-5
