"""CPU oracle for the DistEGNN hot path (FastEGNN forward + virtual-node weighted-mean all-reduce).

TEST INFRASTRUCTURE ONLY.  Nothing in ``distegnn_b200/`` may import this module; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do,
and there only as the checker / the CPU arm — never as the thing measured as "ours" or shipped.

It restates, in plain PyTorch CPU ops and over a ``state_dict`` with the reference's key names,
what ``/root/reference/models/FastEGNN.py`` computes.  Each function cites the reference lines it
follows.  The op sequence (index gathers, ``cat``, dense ``addmm``, ``scatter_add_``) is kept the
same as the reference's so that timing this oracle on host cores is a fair stand-in for timing the
reference's own CPU path (the reference itself cannot travel to the GPU box).

Parity pinning: the reference has no golden vectors of its own (its only test is the unseeded
equivariance script ``equivariant_test.py``).  The oracle is therefore pinned against outputs of the
reference module *itself*, imported unmodified in the build container by ``oracle/make_golden.py``
and committed as fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays them.

Third-party arithmetic restated here: ``torch_geometric.nn.global_mean_pool`` (torch_geometric
2.6.1, requirements.txt:15) = scatter-sum over dim 0 by graph id divided by ``count.clamp(min=1)``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


# --------------------------------------------------------------------------------------------
# small pieces
# --------------------------------------------------------------------------------------------
def segment_sum(data: Tensor, seg: Tensor, num: int) -> Tensor:
    """FastEGNN.py:322-327 — scatter_add_ of [E,K] rows into zeros [num,K]."""
    out = data.new_zeros((num, data.size(1)))
    out.scatter_add_(0, seg.unsqueeze(-1).expand(-1, data.size(1)), data)
    return out


def segment_mean(data: Tensor, seg: Tensor, num: int) -> Tensor:
    """FastEGNN.py:330-337 — sum / count.clamp(min=1); the count is a scatter of a ones tensor of
    the same [E,K] shape (kept so the CPU cost matches)."""
    idx = seg.unsqueeze(-1).expand(-1, data.size(1))
    tot = data.new_zeros((num, data.size(1)))
    cnt = data.new_zeros((num, data.size(1)))
    tot.scatter_add_(0, idx, data)
    cnt.scatter_add_(0, idx, torch.ones_like(data))
    return tot / cnt.clamp(min=1)


def graph_mean_pool(x: Tensor, batch: Tensor, num_graphs: int) -> Tensor:
    """torch_geometric.nn.global_mean_pool (call sites FastEGNN.py:193,222,258): per-graph mean of
    node rows, empty graphs give 0."""
    tot = x.new_zeros((num_graphs, x.size(1)))
    tot.index_add_(0, batch, x)
    cnt = torch.bincount(batch, minlength=num_graphs).clamp(min=1).to(x.dtype)
    return tot / cnt.unsqueeze(-1)


def _lin(sd: StateDict, key: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[key + ".weight"], sd.get(key + ".bias"))


def _mlp2(sd: StateDict, prefix: str, x: Tensor, last_act: bool) -> Tensor:
    """Sequential(Linear, SiLU, Linear[, SiLU]) as declared at FastEGNN.py:69-81,130-141."""
    y = _lin(sd, prefix + ".2", F.silu(_lin(sd, prefix + ".0", x)))
    return F.silu(y) if last_act else y


def _coord_head(sd: StateDict, prefix: str, x: Tensor) -> Tensor:
    """Linear(H,H) → SiLU → Linear(H,1,bias=False) (FastEGNN.py:96-112)."""
    return F.linear(F.silu(_lin(sd, prefix + ".0", x)), sd[prefix + ".2.weight"])


# --------------------------------------------------------------------------------------------
# cross-partition reduction hook
# --------------------------------------------------------------------------------------------
class PartitionReducer:
    """Stands in for ``weighted_average_reduce`` (FastEGNN.py:310-319) when several partitions are
    evaluated inside one process: given each partition's per-graph mean and node count it returns
    Σ_r n_r·mean_r / Σ_r n_r, exactly the arithmetic the reference performs with two all-reduces.
    """

    @staticmethod
    def combine(means: Sequence[Tensor], counts: Sequence[Tensor]) -> Tensor:
        acc = None
        tot = None
        for m, c in zip(means, counts):
            w = c.to(m.dtype).reshape([-1] + [1] * (m.dim() - 1))
            acc = m * w if acc is None else acc + m * w      # data.mul_(weight); all_reduce SUM
            tot = w.clone() if tot is None else tot + w      # all_reduce SUM of the counts
        return acc / tot


# --------------------------------------------------------------------------------------------
# one layer (E_GCL_vel.forward, FastEGNN.py:249-276) evaluated on P >= 1 partitions at once
# --------------------------------------------------------------------------------------------
def _layer(sd: StateDict, pfx: str, parts: List[dict], X: Tensor, Hv: Tensor, C: int,
           normalize: bool, num_graphs: int, trace: Optional[dict]) -> Tuple[Tensor, Tensor]:
    """Runs one E_GCL_vel layer on every partition in ``parts`` (dicts holding h,x,v,row,col,batch,
    edge_attr,node_attr) and returns the new (X, Hv), which are identical on every partition after
    the weighted-mean reductions.  With a single partition no reduction happens, as in the
    reference's ``world_size == 1`` branch."""
    multi = len(parts) > 1
    counts = [torch.bincount(p["batch"], minlength=num_graphs) for p in parts]

    # ---- coord mean (FastEGNN.py:258-261) ----
    cm = [graph_mean_pool(p["x"], p["batch"], num_graphs) for p in parts]
    coord_mean = PartitionReducer.combine(cm, counts) if multi else cm[0]
    # ---- m_X (FastEGNN.py:263-264) ----
    Z = X - coord_mean.unsqueeze(-1)                                   # [B,3,C]
    m_X = torch.einsum("bij,bjk->bik", Z.permute(0, 2, 1), Z)          # [B,C,C]

    per_part = []
    for p in parts:
        h, x, v, row, col, b = p["h"], p["x"], p["v"], p["row"], p["col"], p["batch"]
        # coord2radial, FastEGNN.py:237-246
        dx = x[row] - x[col]
        radial = torch.sum(dx ** 2, 1, keepdim=True)
        if normalize:
            dx = dx / (torch.sqrt(radial).detach() + 1e-8)      # norm detached (FastEGNN.py:243): matters for autograd
        # virtual geometry, FastEGNN.py:252-253
        dX = X[b] - x.unsqueeze(-1)                                    # [N,3,C]
        vr = torch.norm(dX, p=2, dim=1, keepdim=True)                  # [N,1,C]
        # edge_model, FastEGNN.py:144-150
        m = _mlp2(sd, pfx + "edge_mlp", torch.cat([h[row], h[col], radial, p["edge_attr"]], dim=1), True)
        # edge_mode_virtual, FastEGNN.py:154-163
        vin = torch.cat([h.unsqueeze(-1).repeat(1, 1, C), Hv[b], vr, m_X[b]], dim=1)   # [N,2H+1+C,C]
        mv = _mlp2(sd, pfx + "edge_mlp_virtual", vin.permute(0, 2, 1), True)           # [N,C,H]
        # coord_model_vel, FastEGNN.py:166-188 (coords_agg='mean')
        x_new = x + segment_mean(dx * _coord_head(sd, pfx + "coord_mlp_r", m), row, x.size(0))
        phi_xv = _coord_head(sd, pfx + "coord_mlp_r_virtual", mv).permute(0, 2, 1)     # [N,1,C]
        x_new = x_new + torch.mean(-dX * phi_xv, dim=-1)
        x_new = x_new + _mlp2(sd, pfx + "coord_mlp_vel", h, False) * v
        # coord_model_virtual (local mean), FastEGNN.py:191-193
        phi_X = _coord_head(sd, pfx + "coord_mlp_v_virtual", mv).permute(0, 2, 1)      # [N,1,C]
        aggX = graph_mean_pool((dX * phi_X).reshape(x.size(0), -1), b, num_graphs).reshape(-1, 3, C)
        # node_model, FastEGNN.py:203-217
        agg = segment_mean(m, row, h.size(0))
        agg_v = mv.permute(0, 2, 1).mean(dim=-1)                                       # [N,H]
        feats = [h, agg, agg_v] + ([p["node_attr"]] if p["node_attr"] is not None else [])
        h_new = h + _mlp2(sd, pfx + "node_mlp", torch.cat(feats, dim=1), False)
        # node_model_virtual (local mean), FastEGNN.py:220-223
        aggH = graph_mean_pool(mv.permute(0, 2, 1).reshape(h.size(0), -1), b, num_graphs) \
            .reshape(-1, h.size(1), C)
        per_part.append((h_new, x_new, aggX, aggH))

    aggX = PartitionReducer.combine([t[2] for t in per_part], counts) if multi else per_part[0][2]
    aggH = PartitionReducer.combine([t[3] for t in per_part], counts) if multi else per_part[0][3]
    X_new = X + aggX                                                                   # :199
    Hv_new = Hv + _mlp2(sd, pfx + "node_mlp_virtual",
                        torch.cat([Hv, aggH], dim=1).permute(0, 2, 1), False).permute(0, 2, 1)  # :229-233
    for p, t in zip(parts, per_part):
        p["h"], p["x"] = t[0], t[1]
    if trace is not None:
        trace.setdefault("h", []).append([p["h"].clone() for p in parts])
        trace.setdefault("x", []).append([p["x"].clone() for p in parts])
        trace.setdefault("X", []).append(X_new.clone())
        trace.setdefault("Hv", []).append(Hv_new.clone())
    return X_new, Hv_new


# --------------------------------------------------------------------------------------------
# public entry points
# --------------------------------------------------------------------------------------------
def num_layers_of(sd: StateDict) -> int:
    n = 0
    while f"gcl_{n}.edge_mlp.0.weight" in sd:
        n += 1
    return n


def forward_partitions(sd: StateDict, parts_in: Sequence[dict], loc_mean: Tensor, *,
                       normalize: bool = False, trace: Optional[dict] = None
                       ) -> Tuple[List[Tensor], Tensor]:
    """DistEGNN forward over P partitions (one per would-be rank) evaluated in one process.

    ``parts_in[r]`` holds ``node_feat, node_loc, node_vel, edge_index, data_batch, edge_attr,
    node_attr`` for rank r; ``loc_mean`` [B,3] is the global centroid every rank carries
    (distribute_graphs.py:32).  Returns ([node_loc_r for r], virtual_node_loc [B,3,C]).
    Follows FastEGNN.forward (FastEGNN.py:296-307) with ``weighted_average_reduce`` evaluated
    by :class:`PartitionReducer`.
    """
    C = sd["virtual_node_feat"].size(2)
    B = loc_mean.size(0)
    L = num_layers_of(sd)
    Hv = sd["virtual_node_feat"].repeat(B, 1, 1)                  # :299
    X = loc_mean.unsqueeze(-1).repeat(1, 1, C)                    # :300
    parts = []
    for p in parts_in:
        ei = p["edge_index"]
        parts.append(dict(
            h=_lin(sd, "embedding_in", p["node_feat"]),           # :302
            x=p["node_loc"], v=p["node_vel"], row=ei[0], col=ei[1], batch=p["data_batch"],
            edge_attr=p["edge_attr"], node_attr=p.get("node_attr")))
    for i in range(L):
        X, Hv = _layer(sd, f"gcl_{i}.", parts, X, Hv, C, normalize, B, trace)
    return [p["x"] for p in parts], X


def forward(sd: StateDict, node_feat: Tensor, node_loc: Tensor, node_vel: Tensor, loc_mean: Tensor,
            edge_index: Tensor, data_batch: Tensor, edge_attr: Tensor,
            node_attr: Optional[Tensor] = None, *, normalize: bool = False,
            trace: Optional[dict] = None) -> Tuple[Tensor, Tensor]:
    """Single-partition FastEGNN.forward (FastEGNN.py:296-307), world_size == 1."""
    outs, X = forward_partitions(sd, [dict(node_feat=node_feat, node_loc=node_loc, node_vel=node_vel,
                                           edge_index=edge_index, data_batch=data_batch,
                                           edge_attr=edge_attr, node_attr=node_attr)],
                                 loc_mean, normalize=normalize, trace=trace)
    return outs[0], X


def block_diagonal(parts_in: Sequence[dict]) -> dict:
    """Concatenate P partitions of the *same* graphs into one block-diagonal single-process input
    (SURVEY §8c(i)): node arrays concatenated partition after partition but re-sorted so that
    ``data_batch`` stays sorted, edges offset per partition, no cross edges.  Returns the merged
    dict plus ``slices`` to cut the merged output back into partitions."""
    offs, n = [], 0
    for p in parts_in:
        offs.append(n)
        n += p["node_feat"].size(0)
    cat = lambda k: torch.cat([p[k] for p in parts_in], dim=0)
    batch = cat("data_batch")
    order = torch.argsort(batch, stable=True)
    inv = torch.empty_like(order)
    inv[order] = torch.arange(order.numel())
    ei = torch.cat([p["edge_index"] + o for p, o in zip(parts_in, offs)], dim=1)
    merged = dict(node_feat=cat("node_feat")[order], node_loc=cat("node_loc")[order],
                  node_vel=cat("node_vel")[order], data_batch=batch[order], edge_index=inv[ei],
                  edge_attr=cat("edge_attr"),
                  node_attr=None if parts_in[0].get("node_attr") is None else cat("node_attr")[order])
    slices = [inv[o:o + p["node_feat"].size(0)] for p, o in zip(parts_in, offs)]
    return dict(merged=merged, slices=slices)


def init_state_dict(node_feat_nf: int, node_attr_nf: int, edge_attr_nf: int, hidden_nf: int,
                    virtual_channels: int, n_layers: int = 4, seed: int = 0,
                    coord_gain: float = 1e-3, dtype=torch.float32) -> StateDict:
    """A state_dict with the reference's keys/shapes and the reference's init
    distributions (nn.Linear default = kaiming-uniform(a=√5) ⇒ U(±1/√fan_in) for weight and bias;
    xavier-uniform(gain) for the three 1-wide coord heads, FastEGNN.py:97-98).  ``coord_gain`` > 1e-3
    gives the "trained-like" variant used to make coordinate parity tests sensitive."""
    g = torch.Generator().manual_seed(seed)
    H, C = hidden_nf, virtual_channels

    def lin(prefix, fin, fout, bias=True, sd=None):
        bound = 1.0 / (fin ** 0.5)
        sd[prefix + ".weight"] = (torch.rand(fout, fin, generator=g) * 2 - 1) * bound
        if bias:
            sd[prefix + ".bias"] = (torch.rand(fout, generator=g) * 2 - 1) * bound

    def head(prefix, sd):
        lin(prefix + ".0", H, H, sd=sd)
        bound = coord_gain * (6.0 / (H + 1)) ** 0.5
        sd[prefix + ".2.weight"] = (torch.rand(1, H, generator=g) * 2 - 1) * bound

    sd: StateDict = {"virtual_node_feat": torch.randn(1, H, C, generator=g)}
    lin("embedding_in", node_feat_nf, H, sd=sd)
    for i in range(n_layers):
        p = f"gcl_{i}."
        lin(p + "edge_mlp.0", 2 * H + 1 + edge_attr_nf, H, sd=sd)
        lin(p + "edge_mlp.2", H, H, sd=sd)
        lin(p + "edge_mlp_virtual.0", 2 * H + 1 + C, H, sd=sd)
        lin(p + "edge_mlp_virtual.2", H, H, sd=sd)
        head(p + "coord_mlp_r", sd)
        head(p + "coord_mlp_r_virtual", sd)
        head(p + "coord_mlp_v_virtual", sd)
        lin(p + "coord_mlp_vel.0", H, H, sd=sd)
        lin(p + "coord_mlp_vel.2", H, 1, sd=sd)
        lin(p + "node_mlp.0", 3 * H + node_attr_nf, H, sd=sd)
        lin(p + "node_mlp.2", H, H, sd=sd)
        lin(p + "node_mlp_virtual.0", 2 * H, H, sd=sd)
        lin(p + "node_mlp_virtual.2", H, H, sd=sd)
    return {k: v.to(dtype) for k, v in sd.items()}
