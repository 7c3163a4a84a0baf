"""Loss side of the training step: oracle pinned to the reference's own `train_single_epoch` (fixtures
from oracle/make_golden_loss.py), and the fused CUDA loss against both."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import train_loss_oracle as tlo
from tests.helpers import GOLDEN

CASES = ["loss_b1_c5", "loss_b3_c3", "loss_b2_c8_fewnodes"]


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = ast.literal_eval(str(z["meta"]))
    B = len(meta["sizes"])
    smp = [torch.from_numpy(z[f"sample.{i}"]) for i in range(B)]
    return z, meta, smp


@pytest.mark.parametrize("name", CASES)
def test_loss_oracle_matches_reference_train_step(name):
    """float64 replay of utils/train.py:98-147: logged loss and both gradients of the UNMODIFIED train_single_epoch."""
    z, meta, smp = load(name)
    pred = torch.from_numpy(z["pred"]).double().requires_grad_(True)
    V = torch.from_numpy(z["V"]).double().requires_grad_(True)
    loss, logged = tlo.train_loss(pred, torch.from_numpy(z["target"]).double(), V, torch.from_numpy(z["batch"]), smp,
                                  node_counts=[float(len(z["batch"]))], rank=0, sigma=meta["sigma"], weight=meta["weight"],
                                  samples_per_channel=meta["samples"])
    gp, gV = torch.autograd.grad(loss, [pred, V])
    # the reference keeps node counts / the MMD weight / the logged loss in float32 tensors: ~5e-8 relative
    assert abs(float(logged) - float(z["f64.logged"])) <= 1e-7 * float(logged)
    assert float((gp - torch.from_numpy(z["f64.g_pred"])).abs().max()) <= 1e-7 * float(gp.abs().max())
    assert float((gV - torch.from_numpy(z["f64.g_V"])).abs().max()) <= 1e-7 * float(gV.abs().max())


def test_sampling_replays_the_reference_rng_stream():
    """draw_samples consumes the global generator exactly like the reference's per-graph randperm loop."""
    from distegnn_b200.loss import draw_samples
    z, meta, smp = load("loss_b3_c3")
    torch.manual_seed(meta["seed"])
    mine = draw_samples(meta["sizes"], meta["samples"] * meta["C"])
    for i, s in enumerate(smp):
        assert torch.equal(mine[i, :len(s)].long(), s) and bool((mine[i, len(s):] == -1).all())


def _pad(smp, S):
    out = torch.full((len(smp), S), -1, dtype=torch.int32)
    for i, s in enumerate(smp):
        out[i, :len(s)] = s.to(torch.int32)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_fused_loss_against_reference_fixtures(name):
    """CUDA loss (2 launches) vs the reference's float64 results on the same inputs and the same drawn samples."""
    from distegnn_b200 import train_loss
    z, meta, smp = load(name)
    dev = torch.device("cuda:0")
    pred = torch.from_numpy(z["pred"]).to(dev).requires_grad_(True)
    V = torch.from_numpy(z["V"]).to(dev).requires_grad_(True)
    S = meta["samples"] * meta["C"]
    loss, info = train_loss(pred, torch.from_numpy(z["target"]).to(dev), V, torch.from_numpy(z["batch"]).to(dev),
                            world_size=1, mmd_samples=meta["samples"], mmd_sigma=meta["sigma"], mmd_weight=meta["weight"],
                            samples=_pad(smp, S))
    loss.backward()
    torch.cuda.synchronize()
    e_log = abs(float(info["logged"]) - float(z["f64.logged"])) / float(z["f64.logged"])
    e_loss = abs(float(loss) - float(z["f64.oracle_loss"])) / abs(float(z["f64.oracle_loss"]))
    gp, gV = torch.from_numpy(z["f64.g_pred"]), torch.from_numpy(z["f64.g_V"])
    e_gp = float((pred.grad.cpu().double() - gp).abs().max() / gp.abs().max())
    e_gV = float((V.grad.cpu().double() - gV).abs().max() / gV.abs().max())
    print(f"{name}: rel err vs reference fp64  loss {e_loss:.2e}  logged {e_log:.2e}  g_pred {e_gp:.2e}  g_V {e_gV:.2e}")
    assert e_log <= 2e-6 and e_loss <= 2e-6 and e_gp <= 2e-6 and e_gV <= 2e-5


@pytest.mark.gpu
def test_fused_loss_draws_like_the_reference_and_scales_with_accumulation():
    from distegnn_b200 import train_loss
    z, meta, smp = load("loss_b3_c3")
    dev = torch.device("cuda:0")
    args = (torch.from_numpy(z["pred"]).to(dev), torch.from_numpy(z["target"]).to(dev), torch.from_numpy(z["V"]).to(dev),
            torch.from_numpy(z["batch"]).to(dev))
    kw = dict(world_size=1, mmd_samples=meta["samples"], mmd_sigma=meta["sigma"], mmd_weight=meta["weight"])
    torch.manual_seed(meta["seed"])
    l1, info = train_loss(*args, **kw)                      # draws its own samples from the global generator
    assert torch.equal(info["samples"].cpu(), _pad(smp, meta["samples"] * meta["C"]))
    l4, _ = train_loss(*args, accumulation_steps=4, samples=info["samples"], **kw)
    assert abs(float(l1) / 4 - float(l4)) <= 1e-6 * abs(float(l1))


@pytest.mark.gpu
def test_fused_loss_two_rank_weighting_and_folded_collective():
    """world_size = 2 on one GPU: both ranks' partial launches, the packed vectors SUMMED by hand (what the all-reduce
    does), both finalizes — against the oracle's n_r/Σn weighting; loc_mean of both ranks arrives through the same packed
    vector (slot per rank) and a mismatch shows up in out[3]."""
    from distegnn_b200 import _lib
    from distegnn_b200._lib import check, ptr
    from distegnn_b200.loss import graph_offsets
    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    C, samples, sigma, weight, B = 5, 7, 2.0, 0.03, 2
    S = samples * C
    sizes = [[40, 25], [13, 60]]                               # nodes per graph on rank 0 / rank 1
    V = torch.randn(B, 3, C, generator=g)
    ranks = []
    for r in range(2):
        n = sum(sizes[r])
        batch = torch.cat([torch.full((s,), i, dtype=torch.long) for i, s in enumerate(sizes[r])])
        target = torch.randn(n, 3, generator=g) * 2
        pred = target + 0.2 * torch.randn(n, 3, generator=g)
        smp = [torch.randperm(s, generator=g)[:S] for s in sizes[r]]
        ranks.append(dict(batch=batch, target=target, pred=pred, smp=smp, n=n))
    loc_mean = torch.randn(B, 3, generator=g)
    npk = lib.distegnn_loss_packed_floats(B, 2)
    st = torch.cuda.current_stream().cuda_stream
    bufs = []
    for r, d in enumerate(ranks):
        acc, packed = torch.zeros(3, device=dev), torch.zeros(npk, device=dev)
        gV = torch.empty(B, 3, C, device=dev)
        lm = (loc_mean + (1e-3 if r == 1 else 0.0)).to(dev)     # rank 1 disagrees by 1e-3 on purpose
        dd = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
        gp = graph_offsets(dd["batch"], B)
        sm = _pad(d["smp"], S).to(dev)
        check(lib.distegnn_loss_partials(d["n"], B, C, S, 2, r, sigma, ptr(dd["pred"]), ptr(dd["target"]), ptr(V.to(dev)),
                                         ptr(lm), ptr(gp), ptr(sm), ptr(acc), ptr(packed), ptr(gV), st), "partials")
        bufs.append((acc, packed, gV, lm, dd))
    total = bufs[0][1] + bufs[1][1]                              # the all-reduce
    for r, (acc, packed, gV, lm, dd) in enumerate(bufs):
        g_pred, g_Xv, out = torch.empty_like(dd["pred"]), torch.empty(B, 3, C, device=dev), torch.empty(4, device=dev)
        check(lib.distegnn_loss_finalize(ranks[r]["n"], B, C, S, 2, r, sigma, weight, 1, ptr(dd["pred"]), ptr(dd["target"]),
                                         ptr(lm), ptr(acc), ptr(total), ptr(gV), ptr(g_pred), ptr(g_Xv), ptr(out), st),
              "finalize")
        torch.cuda.synchronize()
        p64 = ranks[r]["pred"].double().requires_grad_(True)
        V64 = V.double().requires_grad_(True)
        loss, _ = tlo.train_loss(p64, ranks[r]["target"].double(), V64, ranks[r]["batch"], ranks[r]["smp"],
                                 node_counts=[ranks[0]["n"], ranks[1]["n"]], rank=r, sigma=sigma, weight=weight,
                                 samples_per_channel=samples)
        rp, rV = torch.autograd.grad(loss, [p64, V64])
        logged = sum(tlo.train_loss(ranks[q]["pred"].double(), ranks[q]["target"].double(), V.double(), ranks[q]["batch"],
                                    ranks[q]["smp"], node_counts=[ranks[0]["n"], ranks[1]["n"]], rank=q, sigma=sigma,
                                    weight=weight, samples_per_channel=samples)[1] for q in range(2))
        assert abs(float(out[0]) - float(loss)) <= 2e-6 * abs(float(loss))
        assert abs(float(out[1]) - float(logged)) <= 2e-6 * float(logged)
        assert float((g_pred.cpu().double() - rp).abs().max()) <= 2e-6 * float(rp.abs().max())
        assert float((g_Xv.cpu().double() - rV).abs().max()) <= 2e-5 * float(rV.abs().max())
        assert abs(float(out[3]) - 1e-3) <= 1e-6                 # the loc_mean disagreement is reported
