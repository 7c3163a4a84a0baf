"""Differentiable torch forms of the per-NODE and per-GRAPH dense stages — TEST INFRASTRUCTURE.

Restates what the node / virtual-update / embedding kernels compute (same decomposition: P/Q/Hn split of the first MLP
layers, SUMS instead of means, packed vsum), with the reference lines each stands for.  Until round 2 the backward pass of
the product recomputed these stages here and let torch.autograd (cuBLAS) differentiate them; they are now hand-written
kernels (csrc/node_layer_bwd.cu, csrc/virtual_update.cu), and this file only backs the torch stand-in backend
(tests/shadow_backend.py) those kernels are tested against.  Nothing in the package imports it.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from distegnn_b200 import _lib

Tensor = torch.Tensor
H = _lib.HIDDEN


def field_views(lp: Tensor, A: int, C: int, Na: int) -> Dict[str, Tensor]:
    """Named (differentiable) views into one layer's flat parameter block (layout: distegnn_param_layout)."""
    offs, _ = _lib.param_layout(A, C, Na)
    sizes = dict(E_W1A=(H, H), E_W1B=(H, H), E_B1=(H,), V_W1H=(H, H), V_W1V=(H, H), V_W1M=(C, H), V_B1=(H,),
                 L_W=(H, H), L_B=(H,), L_W3=(H,), L_B3=(1,), N_W1=(3 * H + Na, H), N_B1=(H,), N_W2=(H, H), N_B2=(H,),
                 M_W1=(2 * H, H), M_B1=(H,), M_W2=(H, H), M_B2=(H,))
    out = {}
    for k, shp in sizes.items():
        n = 1
        for s in shp:
            n *= s
        out[k] = lp[offs[k]:offs[k] + n].reshape(shp)
    return out


def projections(h: Tensor, f: Dict[str, Tensor]) -> Tuple[Tensor, Tensor, Tensor]:
    """P, Q, Hn of a layer: the per-node halves of the first edge-MLP / virtual-MLP layers
    (reference FastEGNN.py:69-70, 76-77: Linear(2H+1+A, H) / Linear(2H+1+C, H) applied to a concatenation)."""
    return h @ f["E_W1A"] + f["E_B1"], h @ f["E_W1B"], h @ f["V_W1H"]


def embed_stage(node_feat: Tensor, emb_wt: Tensor, emb_b: Tensor, f0: Dict[str, Tensor]):
    """embedding_in (FastEGNN.py:302) + layer-0 projections -> (h, P, Q, Hn)."""
    h = node_feat @ emb_wt + emb_b
    return (h,) + projections(h, f0)


def node_stage(h: Tensor, x3: Tensor, vel: Tensor, attr: Optional[Tensor], agg_m: Optional[Tensor], agg_x3: Tensor,
               agg_v: Optional[Tensor], trans_v3: Tensor, deg: Tensor, f: Dict[str, Tensor],
               f_next: Optional[Dict[str, Tensor]]):
    """coord_model_vel tail + node_model (FastEGNN.py:177-183, 203-217) + next layer's projections.
    -> (x3', h', P', Q', Hn')  (the last four are None for the last layer: f_next is None)."""
    phiv = F.silu(h @ f["L_W"] + f["L_B"]) @ f["L_W3"] + f["L_B3"]
    xn = x3 + agg_x3 / deg + trans_v3 + phiv.unsqueeze(1) * vel
    if f_next is None:
        return xn, None, None, None, None
    cat = [h, agg_m / deg, agg_v] + ([attr] if attr is not None else [])
    hn = h + F.silu(torch.cat(cat, 1) @ f["N_W1"] + f["N_B1"]) @ f["N_W2"] + f["N_B2"]
    return (xn, hn) + projections(hn, f_next)


def virtual_update_stage(vsum: Tensor, Xv: Tensor, Hv: Tensor, f: Optional[Dict[str, Tensor]],
                         f_next: Optional[Dict[str, Tensor]], init: bool, C: int):
    """coord_model_virtual tail + node_model_virtual + next layer's virtual geometry (FastEGNN.py:193-199, 222-234,
    258-264) from the packed, all-reduced statistics.  -> (Xv', Hv', G')  (Hv', G' None for the last layer)."""
    B = vsum.shape[0]
    n = vsum[:, 3].detach().clamp(min=1)
    if not init:
        Xv = Xv + vsum[:, 4:4 + 3 * C].reshape(B, 3, C) / n.view(B, 1, 1)
    if f_next is None:
        return Xv, None, None
    if not init:
        agg = vsum[:, 4 + 3 * C:].reshape(B, C, H) / n.view(B, 1, 1)
        Hv = Hv + F.silu(torch.cat([Hv, agg], -1) @ f["M_W1"] + f["M_B1"]) @ f["M_W2"] + f["M_B2"]
    xbar = vsum[:, 0:3] / n.view(B, 1)
    Z = Xv - xbar.unsqueeze(-1)
    mX = torch.einsum("bdi,bdj->bij", Z, Z)
    G = Hv @ f_next["V_W1V"] + torch.einsum("bjc,jn->bcn", mX, f_next["V_W1M"]) + f_next["V_B1"]
    return Xv, Hv, G
