"""Input wire format on CPU: shard round trip, and FastEGNN.forward fed with the shard's pre-sorted
CSRGraph gives the same outputs and gradients as with the int64 edge_index (kernels replaced by the torch stand-in)."""
import numpy as np
import pytest
import torch

from distegnn_b200 import FastEGNN
from distegnn_b200.shards import CSRGraph, Shard, read_shard, shard_from_forward_inputs, write_shard
from tests.helpers import SINGLE_CASES, golden_inputs, load_golden
from tests.shadow_backend import ShadowBackend


def test_shard_round_trip(tmp_path):
    z, kw, sd = load_golden("fluid160_c5")
    inp = golden_inputs(z)
    arrays = shard_from_forward_inputs(inp, target=inp["node_loc"] + 1)
    p = str(tmp_path / "g0_rank0.shard")
    write_shard(p, arrays)
    sh = read_shard(p)
    assert set(sh.t) == set(arrays)
    for k, v in arrays.items():
        assert sh.t[k].dtype == torch.from_numpy(v).dtype and tuple(sh.t[k].shape) == v.shape
        assert np.array_equal(sh.t[k].numpy(), v), k
    # CSR really is the graph: same multiset of (row, col, attr) as the COO input
    kwargs = sh.to("cpu", non_blocking=False)
    g = kwargs["edge_index"]
    assert isinstance(g, CSRGraph) and g.num_nodes == inp["node_loc"].shape[0] and g.num_edges == inp["edge_index"].shape[1]
    ei = g.edge_index()
    assert bool((ei[0][1:] >= ei[0][:-1]).all())
    key = lambda e, a: sorted(zip(e[0].tolist(), e[1].tolist(), [tuple(r) for r in a.tolist()]))
    assert key(ei, kwargs["edge_attr"]) == key(inp["edge_index"], inp["edge_attr"])
    with pytest.raises(ValueError):
        (tmp_path / "bad").write_bytes(b"not a shard at all........")
        read_shard(str(tmp_path / "bad"))


@pytest.mark.parametrize("name", SINGLE_CASES)
def test_forward_and_gradients_from_a_shard_match_the_edge_index_path(name, tmp_path):
    z, kw, sd = load_golden(name)
    inp = golden_inputs(z)
    p = str(tmp_path / "s.shard")
    write_shard(p, shard_from_forward_inputs(inp))
    kwargs = read_shard(p).to("cpu", non_blocking=False)
    outs = []
    for args in (inp, kwargs):
        m = FastEGNN(hidden_nf=64, world_size=1, **kw)
        m.load_state_dict(sd)
        m._backend = ShadowBackend()
        out, X = m(**args)
        (out.square().sum() + X.square().sum()).backward()
        outs.append((out.detach(), X.detach(), {k: p_.grad.clone() for k, p_ in m.named_parameters() if p_.grad is not None}))
    # same graph, same arithmetic; only the order of equal-row edges may differ (stable sort keeps it) -> tight tolerance
    assert float((outs[0][0] - outs[1][0]).abs().max()) <= 1e-6
    assert float((outs[0][1] - outs[1][1]).abs().max()) <= 1e-6
    for k, g in outs[0][2].items():
        den = float(g.abs().max())
        if den > 0:
            assert float((g - outs[1][2][k]).abs().max()) <= 1e-4 * den, k
