"""Loss side of the reference's training step (utils/train.py:98-147), fused.

    loss, info = distegnn_b200.train_loss(loc_pred, loc_target, virtual_node_loc, batch, world_size=..., mmd_samples=50,
                                          mmd_sigma=3, mmd_weight=0.01, accumulation_steps=4, loc_mean=loc_mean)
    loss.backward()

computes what the reference computes between the model call and `loss_loc.backward()`:

  * node-count weighted MSE  `world_size · n_r/Σn · MSE(loc_pred, loc_target)`                        (train.py:98-110)
  * the MMD regulariser between the virtual coordinates and `mmd_samples·C` sampled target positions per graph, kernel
    `exp(−‖x−y‖/(2σ²))`                                                                               (train.py:11-14, 119-147)
  * the per-step collectives — total node count (:104), logged loss (:109), `loc_mean` consistency check (:52-61) —
    folded into ONE packed SUM all-reduce

in two kernel launches + one collective (csrc/loss.cu) instead of a Python loop over the graphs with ~20 ATen launches
each and three collectives; the forward already produces the gradients w.r.t. `loc_pred` and `virtual_node_loc`, the
backward only scales them by the incoming gradient.  Sampling follows the reference exactly — one
`torch.randperm(num_node)[:S]` per graph on the global CPU generator, in graph order — unless `samples` is passed in.
CUDA only (no CPU path); `info` holds device scalars (`logged`, `mmd`, `loc_mean_dev`) — reading them synchronises.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import check, ptr

Tensor = torch.Tensor


def graph_offsets(batch: Tensor, n_graphs: int) -> Tensor:
    """[B+1] int64 first-node offsets of a sorted `batch` vector (device, no host sync)."""
    return torch.searchsorted(batch.contiguous(), torch.arange(n_graphs + 1, device=batch.device, dtype=batch.dtype))


def draw_samples(node_counts: Sequence[int], num_sample: int) -> Tensor:
    """The reference's sampling (train.py:124-129): `torch.randperm(num_node)[:num_sample]` per graph, in graph order, on
    the global CPU generator → int32 [B, num_sample] of graph-local indices, −1 where a graph has fewer nodes."""
    out = torch.full((len(node_counts), num_sample), -1, dtype=torch.int32)
    for i, n in enumerate(node_counts):
        idx = torch.randperm(int(n))[:num_sample]
        out[i, :idx.numel()] = idx.to(torch.int32)
    return out


class _TrainLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, Xv, target, gptr, samples, loc_mean, cfg):
        lib = _lib.load()
        dev = pred.device
        N, (B, _, Cn) = int(pred.shape[0]), Xv.shape
        S, world, rank = int(samples.shape[1]), cfg["world"], cfg["rank"]
        stream = _lib.stream_ptr(dev)
        npk = lib.distegnn_loss_packed_floats(B, world)
        scratch = torch.zeros(3 + npk, dtype=torch.float32, device=dev)          # acc[3] | packed[npk]
        acc, packed = scratch[:3], scratch[3:]
        gV_raw = torch.empty(B, 3, Cn, dtype=torch.float32, device=dev)
        p, t, xv = pred.detach().contiguous(), target.contiguous(), Xv.detach().contiguous()
        with torch.cuda.device(dev):
            check(lib.distegnn_loss_partials(N, B, Cn, S, world, rank, cfg["sigma"], ptr(p), ptr(t), ptr(xv), ptr(loc_mean),
                                             ptr(gptr), ptr(samples), ptr(acc), ptr(packed), ptr(gV_raw), stream),
                  "loss_partials")
            if world > 1:                                    # the ONE collective of the step
                comm, be = cfg.get("comm"), cfg.get("backend")
                if comm is not None and be is not None and npk <= comm.max_slots * comm.slot_floats:
                    be.allreduce_packed(comm, packed)
                else:
                    import torch.distributed as dist
                    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=cfg.get("group"))
            g_pred = torch.empty_like(p)
            g_Xv = torch.empty_like(xv)
            out = torch.empty(4, dtype=torch.float32, device=dev)
            check(lib.distegnn_loss_finalize(N, B, Cn, S, world, rank, cfg["sigma"], cfg["weight"], cfg["accum"], ptr(p),
                                             ptr(t), ptr(loc_mean), ptr(acc), ptr(packed), ptr(gV_raw), ptr(g_pred),
                                             ptr(g_Xv), ptr(out), stream), "loss_finalize")
        ctx.save_for_backward(g_pred, g_Xv)
        ctx.mark_non_differentiable(out)
        return out[0].clone(), out

    @staticmethod
    def backward(ctx, g_loss, _g_out):
        g_pred, g_Xv = ctx.saved_tensors
        return g_loss * g_pred, g_loss * g_Xv, None, None, None, None, None


def train_loss(loc_pred: Tensor, loc_target: Tensor, virtual_node_loc: Tensor, batch: Tensor, *, world_size: int = 1,
               mmd_samples: int = 50, mmd_sigma: float = 3.0, mmd_weight: float = 0.01, accumulation_steps: int = 1,
               loc_mean: Optional[Tensor] = None, samples: Optional[Tensor] = None,
               node_counts: Optional[Sequence[int]] = None, process_group=None, model=None
               ) -> Tuple[Tensor, Dict[str, Tensor]]:
    """See the module docstring.  `node_counts` (host ints per graph, e.g. from the loader's `ptr`) avoids the one host
    sync needed to size the reference's `randperm` draws; `samples` (int32 [B,S], graph-local, −1 padded) overrides the
    draw; `model` (a distegnn_b200.FastEGNN of a multi-partition job) lends its peer-memory communicator to the
    collective."""
    if loc_pred.device.type != "cuda":
        raise _lib.DistEGNNError("distegnn_b200.train_loss runs only on CUDA tensors (no CPU path)")
    if loc_pred.shape != loc_target.shape or loc_pred.dim() != 2 or loc_pred.shape[1] != 3:
        raise ValueError("loc_pred / loc_target must both be [N,3]")
    if virtual_node_loc.dim() != 3 or virtual_node_loc.shape[1] != 3:
        raise ValueError("virtual_node_loc must be [B,3,C]")
    dev = loc_pred.device
    B, _, Cn = virtual_node_loc.shape
    if Cn > _lib.MAX_CHANNELS:
        raise ValueError(f"virtual_channels={Cn} exceeds the compiled limit {_lib.MAX_CHANNELS}")
    S = int(mmd_samples) * Cn                                      # train.py:122
    gptr = graph_offsets(batch, B)
    if samples is None:
        if node_counts is None:
            node_counts = (gptr[1:] - gptr[:-1]).tolist()          # one host sync (the reference has B of them)
        samples = draw_samples(node_counts, S)
    samples = samples.to(device=dev, dtype=torch.int32).contiguous()
    if samples.shape != (B, S):
        raise ValueError(f"samples must be [B={B}, S={S}]")
    rank = 0
    if world_size > 1:
        import torch.distributed as dist
        rank = dist.get_rank(process_group)
    cfg = dict(world=int(world_size), rank=rank, sigma=float(mmd_sigma), weight=float(mmd_weight),
               accum=int(accumulation_steps), group=process_group)
    if model is not None and getattr(model, "_comm", None):
        from .backend import cuda_backend
        cfg["comm"], cfg["backend"] = model._comm, cuda_backend()
    lm = None if loc_mean is None else loc_mean.detach().to(torch.float32).contiguous()
    loss, out = _TrainLoss.apply(loc_pred.to(torch.float32), virtual_node_loc.to(torch.float32),
                                 loc_target.detach().to(torch.float32), gptr, samples, lm, cfg)
    return loss, {"logged": out[1], "mmd": out[2], "loc_mean_dev": out[3], "samples": samples}
