"""The oracle (oracle/fastegnn_oracle.py) against fixtures produced by the unmodified reference
module (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import fastegnn_oracle as orc
from tests.helpers import (DIST_CASE, SINGLE_CASES, golden_inputs, golden_trace, load_golden, max_abs)

# The oracle replays the reference's op sequence, so on the same machine/torch it is bit-identical;
# the tolerance only allows for a different BLAS blocking on another host.
TOL = 2e-6


@pytest.mark.parametrize("name", SINGLE_CASES)
def test_oracle_matches_reference_fp32(name):
    z, kw, sd = load_golden(name)
    inp = golden_inputs(z)
    tr = {}
    out, X = orc.forward(sd, **inp, normalize=kw["normalize"], trace=tr)
    assert max_abs(out, torch.from_numpy(z["out.node_loc"])) <= TOL
    assert max_abs(X, torch.from_numpy(z["out.virtual_loc"])) <= TOL
    for key in ("h", "x"):
        for mine, ref in zip(tr[key], golden_trace(z, key)):
            assert max_abs(mine[0], ref) <= 5e-6 * max(1.0, float(ref.abs().max()))
    for key in ("Hv", "X"):
        for mine, ref in zip(tr[key], golden_trace(z, key)):
            assert max_abs(mine, ref) <= 5e-6 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("name", SINGLE_CASES)
def test_oracle_fp64_matches_reference_fp64(name):
    z, kw, sd = load_golden(name)
    inp = golden_inputs(z)
    sd64 = {k: v.double() for k, v in sd.items()}
    inp64 = {k: (v.double() if (v is not None and v.is_floating_point()) else v) for k, v in inp.items()}
    out, X = orc.forward(sd64, **inp64, normalize=kw["normalize"])
    assert max_abs(out, torch.from_numpy(z["out64.node_loc"])) <= 1e-12
    assert max_abs(X, torch.from_numpy(z["out64.virtual_loc"])) <= 1e-12


def _dist_parts(z):
    return [golden_inputs(z, f"in{r}.") for r in range(2)]


def test_oracle_partitions_match_reference_world_size_2():
    """forward_partitions == the reference's real world_size=2 branch run under gloo."""
    z, kw, sd = load_golden(DIST_CASE)
    parts = _dist_parts(z)
    outs, X = orc.forward_partitions(sd, parts, parts[0]["loc_mean"], normalize=kw["normalize"])
    for r in range(2):
        assert max_abs(outs[r], torch.from_numpy(z[f"out{r}.node_loc"])) <= TOL
        assert max_abs(X, torch.from_numpy(z[f"out{r}.virtual_loc"])) <= TOL


def test_block_diagonal_equivalence():
    """DistEGNN on P partitions ≡ one block-diagonal graph."""
    z, kw, sd = load_golden(DIST_CASE)
    parts = _dist_parts(z)
    bd = orc.block_diagonal(parts)
    m = bd["merged"]
    out, X = orc.forward(sd, m["node_feat"], m["node_loc"], m["node_vel"], parts[0]["loc_mean"],
                         m["edge_index"], m["data_batch"], m["edge_attr"], m["node_attr"],
                         normalize=kw["normalize"])
    for r in range(2):
        assert max_abs(out[bd["slices"][r]], torch.from_numpy(z[f"out{r}.node_loc"])) <= 5e-6
    assert max_abs(X, torch.from_numpy(z["out0.virtual_loc"])) <= 5e-6


def _rotation(seed):
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return torch.from_numpy(q.astype(np.float32))


def test_oracle_equivariance_like_reference_script():
    """equivariant_test.py:12-62 restated with seeds: 10 nodes, 20 random edges (self loops and
    duplicates allowed), F=1, A=1, C=3, atol 1e-4."""
    sd = orc.init_state_dict(1, 0, 1, 64, 3, n_layers=4, seed=7)
    g = torch.Generator().manual_seed(11)
    n, e = 10, 20
    x = torch.rand(n, 3, generator=g) * 10
    v = torch.rand(n, 3, generator=g) * 10
    f = torch.rand(n, 1, generator=g) * 10
    ei = torch.randint(0, 10, (2, e), generator=g)
    ea = torch.rand(e, 1, generator=g) * 10
    b = torch.zeros(n, dtype=torch.long)
    R, t = _rotation(3), torch.randn(3, generator=g) * 5
    out, _ = orc.forward(sd, f, x, v, x.mean(0, keepdim=True), ei, b, ea)
    xr = x @ R + t
    out_r, _ = orc.forward(sd, f, xr, v @ R, xr.mean(0, keepdim=True), ei, b, ea)
    assert torch.allclose(out @ R + t, out_r, atol=1e-4)


def test_init_state_dict_keys_match_reference():
    z, kw, sd = load_golden("fluid160_c5")
    mine = orc.init_state_dict(kw["node_feat_nf"], kw["node_attr_nf"], kw["edge_attr_nf"], 64,
                               kw["virtual_channels"], kw["n_layers"])
    assert {k: tuple(v.shape) for k, v in mine.items()} == {k: tuple(v.shape) for k, v in sd.items()}
