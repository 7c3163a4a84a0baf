"""Multi-GPU parity check of the CUDA path against the partitioned oracle — TEST INFRASTRUCTURE (the checker).

Called under torchrun (one rank per GPU) by ``bench.py`` before its timed region, by ``scripts/dist_parity.py`` and by
the ``-m gpu`` multi-device test.  Every rank builds the same seeded graph, takes its partition (the reference's
partitioners restated in ``distegnn_b200/synth.py``: datasets/distribute_graphs.py:17-51 random, :118-143 k-means), runs
``FastEGNN(world_size=N)`` through the CUDA kernels with the real cross-rank exchange, and rank 0 compares all ranks'
outputs with ``fastegnn_oracle.forward_partitions`` in float64 — which is pinned to the reference's own
``world_size=2`` run (tests/test_oracle_golden.py) — on the SAME partitions (models/FastEGNN.py:195-197, 225-227,
259-261, 310-319).

Gates (the single-GPU ones): |out − ref64| ≤ 1e-5·max(1,|ref|), relative displacement error ≤ 1e-4, virtual coordinates
within 1e-5 and BIT-IDENTICAL on all ranks.  With ``grads=True`` the parameter gradients of Σ_r <out_r, cot_r> +
<X, cot_X>, summed over the ranks, are compared with float64 autograd through the oracle; each parameter is gated at
max(2e-4, 3 × the oracle's own fp32-vs-fp64 difference on that parameter) relative to the gradient's max-abs.
"""
from __future__ import annotations

import time
from typing import Dict, Optional

import torch
import torch.distributed as dist


def _gather_rows(t: torch.Tensor, sizes, dev):
    mx = max(sizes)
    pad = torch.zeros(mx, t.shape[1], device=dev)
    pad[:t.shape[0]] = t
    bufs = [torch.empty(mx, t.shape[1], device=dev) for _ in sizes]
    dist.all_gather(bufs, pad)
    return [b[:n] for b, n in zip(bufs, sizes)]


def check_case(workload: str, n_nodes: Optional[int], split_mode: str, dev: torch.device, *, grads: bool = False,
               seed: int = 11, cuda_graph: bool = False, coord_gain: float = 0.05) -> Dict:
    """Run one case on all ranks; every rank returns the same dict (``pass`` is broadcast from rank 0)."""
    from distegnn_b200 import FastEGNN, synth
    from oracle import fastegnn_oracle as orc
    rank, world = dist.get_rank(), dist.get_world_size()
    w = synth.WORKLOADS[workload]
    t0 = time.perf_counter()
    parts = synth.make_partitions(w, world_size=world, split_mode=split_mode, seed=seed, n_nodes=n_nodes,
                                  only_rank=None if rank == 0 else rank)
    sd = orc.init_state_dict(w.node_feat_nf, w.node_attr_nf, w.edge_attr_nf, 64, w.virtual_channels, 4, seed=3,
                             coord_gain=coord_gain)
    m = FastEGNN(node_feat_nf=w.node_feat_nf, node_attr_nf=w.node_attr_nf, edge_attr_nf=w.edge_attr_nf,
                 hidden_nf=64, virtual_channels=w.virtual_channels, world_size=world, n_layers=4,
                 normalize=w.normalize)
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    m.cuda_graph = cuda_graph
    inp = {k: (v.to(dev) if v is not None else None) for k, v in parts[rank].items()}
    with torch.no_grad():
        out, X = m(**inp)
        if cuda_graph:                                   # second call = a graph REPLAY, must give the same numbers
            out, X = m(**inp)
    torch.cuda.synchronize()
    n_mine = torch.tensor([out.shape[0], inp["edge_index"].shape[1]], device=dev)
    n_all = [torch.zeros_like(n_mine) for _ in range(world)]
    dist.all_gather(n_all, n_mine)
    sizes = [int(t[0]) for t in n_all]
    edges = [int(t[1]) for t in n_all]
    outs = _gather_rows(out, sizes, dev)
    Xs = [torch.empty_like(X) for _ in range(world)]
    dist.all_gather(Xs, X)
    res = {"workload": workload, "nodes": sum(sizes), "edges_sum_p": sum(edges), "split_mode": split_mode,
           "world": world, "collective": "p2p-fused" if getattr(m, "_comm", None) else "torch.distributed",
           "cuda_graph": bool(cuda_graph)}
    ok = True
    if rank == 0:
        sd64 = {k: v.double() for k, v in sd.items()}
        p64 = [{k: (v.double() if (v is not None and v.is_floating_point()) else v) for k, v in p.items()} for p in parts]
        refs, refX = orc.forward_partitions(sd64, [{k: v for k, v in p.items() if k != "loc_mean"} for p in p64],
                                            p64[0]["loc_mean"], normalize=w.normalize)
        worst_abs = worst_rel = worst_x = 0.0
        same = True
        for r in range(world):
            o = outs[r].cpu().double()
            err = float((o - refs[r]).abs().max()) if o.numel() else 0.0
            disp = float((refs[r] - p64[r]["node_loc"]).abs().max()) if o.numel() else 1.0
            ex = float((Xs[r].cpu().double() - refX).abs().max())
            same &= bool((Xs[r] == Xs[0]).all())
            ok &= err <= 1e-5 * max(1.0, float(refs[r].abs().max()) if o.numel() else 1.0) and err / disp <= 1e-4
            ok &= ex <= 1e-5
            worst_abs, worst_rel, worst_x = max(worst_abs, err), max(worst_rel, err / disp), max(worst_x, ex)
        ok &= same
        res.update({"abs": worst_abs, "rel_disp": worst_rel, "virtual": worst_x, "bit_identical_across_ranks": same})
    if grads:
        g = torch.Generator().manual_seed(17)
        cots = [torch.randn(n, 3, generator=g) for n in sizes]
        cotX = torch.randn(X.shape, generator=g)
        m.train()
        m.cuda_graph = False
        out, X = m(**inp)
        ((out * cots[rank].to(dev)).sum() + (X * cotX.to(dev)).sum()).backward()
        names = [k for k, _ in m.named_parameters()]
        flat = torch.cat([(p_.grad if p_.grad is not None else torch.zeros_like(p_)).reshape(-1)
                          for _, p_ in m.named_parameters()])
        dist.all_reduce(flat)                                    # Σ over ranks of each rank's parameter gradient
        if rank == 0:
            def oracle_grads(dtype):
                sdx = {k: v.to(dtype).requires_grad_(True) for k, v in sd.items()}
                px = [{k: (v.to(dtype) if (v is not None and v.is_floating_point()) else v) for k, v in p.items()}
                      for p in parts]
                rs, rX = orc.forward_partitions(sdx, [{k: v for k, v in p.items() if k != "loc_mean"} for p in px],
                                                px[0]["loc_mean"], normalize=w.normalize)
                loss = sum((rs[r] * cots[r].to(dtype)).sum() for r in range(world)) + world * (rX * cotX.to(dtype)).sum()
                return torch.autograd.grad(loss, [sdx[k] for k in names], allow_unused=True)
            rg, rg32 = oracle_grads(torch.float64), oracle_grads(torch.float32)
            off, worst, wk, wtol = 0, 0.0, "", 0.0
            for k, r_, q_ in zip(names, rg, rg32):
                n = sd[k].numel()
                mine = flat[off:off + n].cpu().double().reshape(sd[k].shape)
                off += n
                if r_ is None or float(r_.abs().max()) == 0.0:
                    ok &= float(mine.abs().max()) == 0.0
                    continue
                den = float(r_.abs().max())
                e = float((mine - r_).abs().max() / den)
                noise = float((q_.double() - r_).abs().max() / den)
                tol = max(2e-4, 3.0 * noise)
                ok &= e <= tol
                if e / tol > (worst / wtol if wtol else -1.0):
                    worst, wk, wtol = e, k, tol
            res.update({"grads_worst": worst, "grads_worst_param": wk, "grads_gate": wtol})
    comm = getattr(m, "_comm", None)
    if comm:
        st = comm.status()                               # != 0: an exchange timed out waiting for a peer
        res["comm_status"] = st
        ok &= st == 0
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)          # rank 0 holds the numeric verdict, every rank its comm status
    res["pass"] = bool(flag.item())
    res["seconds"] = round(time.perf_counter() - t0, 2)
    m.release_comm()                                     # collective teardown of the peer mappings
    return res
