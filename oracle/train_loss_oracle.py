"""CPU restatement of the loss side of the reference's training step (utils/train.py:98-147) — TEST INFRASTRUCTURE.

    node-count weighted MSE  (train.py:98-110)   loss = world_size · n_r/Σn · MSE(loc_pred, loc_target)
    logged loss              (train.py:106-108)  Σ_r n_r/Σn · MSE_r  (what `result['loss']` accumulates, × batch_size)
    MMD regulariser          (train.py:119-147)  per graph i: S = samples·C target positions drawn with
                                                 `torch.randperm(num_node)[:S]`, k(x,y) = exp(−‖x−y‖₂ / (2σ²)) (:11-14, the
                                                 distance is NOT squared), l_vv = Σk(V,V)/B/C², l_rv = 2Σk(R,V)/B/S/C,
                                                 loss += weight · world_size · n_r/Σn · (l_vv − l_rv)

Pinned by tests/golden/loss_*.npz, which oracle/make_golden_loss.py produces by driving the UNMODIFIED
`train_single_epoch` for one optimisation step (SGD, lr 1, so the parameter change IS the gradient).
The product (distegnn_b200/loss.py + csrc/loss.cu) is tested against this file; nothing in the package imports it.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor


def mmd_kernel(x: Tensor, y: Tensor, sigma: float) -> Tensor:
    """utils/train.py:11-14."""
    return torch.exp(-torch.cdist(x, y, p=2) / (2 * sigma * sigma))


def draw_samples(batch: Tensor, n_graphs: int, num_sample: int) -> List[Tensor]:
    """The reference's sampling (train.py:124-133): one `torch.randperm(num_node)[:num_sample]` per graph, in graph
    order, on the global CPU generator.  Returns per-graph LOCAL indices (into `loc_target[batch == i]`)."""
    out = []
    for i in range(n_graphs):
        num_node = int((batch == i).sum())
        out.append(torch.randperm(num_node)[:num_sample])
    return out


def train_loss(loc_pred: Tensor, loc_target: Tensor, virtual_node_loc: Tensor, batch: Tensor,
               samples: Sequence[Tensor], *, node_counts: Sequence[float], rank: int, sigma: float, weight: float,
               samples_per_channel: int, accumulation_steps: int = 1) -> Tuple[Tensor, Tensor]:
    """(loss to back-propagate on this rank, this rank's term of the logged loss).  `node_counts[r]` = nodes on rank r
    (the reference all-reduces them, train.py:100-104); `samples[i]` = local indices drawn for graph i."""
    world = len(node_counts)
    n_r, n_tot = float(node_counts[rank]), float(sum(node_counts))
    B, _, C = virtual_node_loc.shape
    mse = torch.nn.functional.mse_loss(loc_pred, loc_target)
    loss_loc = n_r / n_tot * mse
    logged = loss_loc.detach().clone()
    loss = world * loss_loc
    V = virtual_node_loc.permute(0, 2, 1)
    num_sample = samples_per_channel * C
    l_vv = loc_pred.new_zeros(())
    l_rv = loc_pred.new_zeros(())
    for i in range(B):
        tgt_i = loc_target[batch == i][samples[i]]
        l_vv = l_vv + mmd_kernel(V[i], V[i], sigma).sum()
        l_rv = l_rv + mmd_kernel(tgt_i, V[i], sigma).sum()
    l_vv = l_vv / B / C / C
    l_rv = 2 * l_rv / B / num_sample / C
    loss = loss + weight * world * n_r / n_tot * (l_vv - l_rv)
    return loss / float(accumulation_steps), logged
