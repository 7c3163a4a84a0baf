"""Golden vectors for the loss side of the training step, from the UNMODIFIED reference.

    python oracle/make_golden_loss.py         # build container only (/root/reference)

Drives `utils/train.py:train_single_epoch` as it lies for ONE optimisation step on a one-element loader with a stand-in
model whose parameters ARE the prediction tensors (`loc_pred`, `virtual_node_loc`), SGD with lr = 1 and
accumulation_steps = 1: the parameter change is then exactly the gradient of the reference's loss (node-count weighted
MSE + MMD), and the function's return value is its logged loss.  (In fp32 the subtraction old − new only
resolves the gradient to ~1e-3 relative; the fp64 run is the one the tests pin gradients to.)  Patches, none of which touch arithmetic: `tqdm` ->
pass-through, `torch.tensor(..., device='cuda')` -> CPU (train.py:100,102 hard-code the device).  The RNG state right
before the call is the `torch.manual_seed` below, so `torch.randperm` inside the reference is reproducible and the drawn
indices are stored with the fixture.  Test infrastructure; writes tests/golden/loss_*.npz.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


class _Bar:
    def __init__(self, it):
        self.it = it

    def __iter__(self):
        return iter(self.it)

    def set_postfix(self, *_a, **_k):
        pass


def import_train():
    sys.path.insert(0, REF)
    import utils.train as T          # the unmodified reference module
    T.tqdm = lambda it, **kw: _Bar(it)
    return T


class FastEGNN(nn.Module):           # the reference dispatches on this class name (train.py:19-21,64,119)
    def __init__(self, loc_pred, virtual_loc):
        super().__init__()
        self.node_attr_nf = 0
        self.loc_pred = nn.Parameter(loc_pred.clone())
        self.virtual_loc = nn.Parameter(virtual_loc.clone())

    def forward(self, *args):
        return self.loc_pred, self.virtual_loc


def make_case(seed, sizes, C, samples, sigma, weight, dtype):
    # the SAME fp32 numbers for both precisions (the fp64 run is the tolerance basis of the fp32 product)
    g = torch.Generator().manual_seed(seed)
    n = sum(sizes)
    batch = torch.cat([torch.full((s,), i, dtype=torch.long) for i, s in enumerate(sizes)])
    target = torch.randn(n, 3, generator=g) * 2.0
    pred = (target + 0.1 * torch.randn(n, 3, generator=g)).to(dtype)
    target = target.to(dtype)
    B = len(sizes)
    V = torch.randn(B, 3, C, generator=g).to(dtype)
    V[0, :, 0] = V[0, :, 1] if C > 1 else V[0, :, 0]           # a coincident pair: the d = 0 corner of cdist's gradient
    cfg = types.SimpleNamespace(mmd=types.SimpleNamespace(samples=samples, sigma=sigma, weight=weight),
                                accumulation_steps=1)
    return dict(batch=batch, target=target, pred=pred, V=V, cfg=cfg, C=C, sizes=sizes)


def run_reference(T, case, seed):
    c = case
    data = types.SimpleNamespace(x=torch.zeros(len(c["batch"]), 1), pos=c["target"].clone(), vel=torch.zeros_like(c["target"]),
                                 attr=torch.zeros(len(c["batch"]), 1), batch=c["batch"], loc_mean=torch.zeros(len(c["sizes"]), 3),
                                 target=c["target"], edge_index=torch.zeros(2, 0, dtype=torch.long),
                                 edge_attr=torch.zeros(0, 2))
    model = FastEGNN(c["pred"], c["V"])
    opt = torch.optim.SGD(model.parameters(), lr=1.0)
    real_tensor = torch.tensor
    torch.tensor = lambda *a, **k: real_tensor(*a, **{kk: vv for kk, vv in k.items() if kk != "device"})
    try:
        torch.manual_seed(seed)
        logged = T.train_single_epoch("cpu", model, [data], opt, None, nn.MSELoss(), "nbody", c["cfg"], 0, "train",
                                      c["C"], 1)
    finally:
        torch.tensor = real_tensor
    g_pred = c["pred"] - model.loc_pred.detach()             # lr = 1: old − new = gradient
    g_V = c["V"] - model.virtual_loc.detach()
    return logged, g_pred, g_V


def main():
    os.makedirs(OUT, exist_ok=True)
    T = import_train()
    sys.path.insert(0, ROOT)
    from oracle import train_loss_oracle as tlo
    cases = {"loss_b1_c5": (3, [400], 5, 50, 3.0, 0.01), "loss_b3_c3": (4, [60, 7, 130], 3, 10, 1.5, 0.05),
             "loss_b2_c8_fewnodes": (5, [9, 300], 8, 50, 3.0, 0.01)}
    for name, (seed, sizes, C, samples, sigma, weight) in cases.items():
        blob = {}
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            c = make_case(seed, sizes, C, samples, sigma, weight, dtype)
            logged, g_pred, g_V = run_reference(T, c, seed)
            # replay the sampling the reference just did, to store the indices next to its results
            torch.manual_seed(seed)
            smp = tlo.draw_samples(c["batch"], len(sizes), samples * C)
            loss, lg = tlo.train_loss(c["pred"].clone().requires_grad_(True), c["target"], c["V"], c["batch"], smp,
                                      node_counts=[float(len(c["batch"]))], rank=0, sigma=sigma, weight=weight,
                                      samples_per_channel=samples)
            blob.update({f"{tag}.logged": np.array(logged), f"{tag}.g_pred": g_pred.numpy(), f"{tag}.g_V": g_V.numpy(),
                         f"{tag}.oracle_loss": np.array(float(loss))})
            if tag == "f32":
                blob.update(batch=c["batch"].numpy(), target=c["target"].numpy(), pred=c["pred"].numpy(), V=c["V"].numpy(),
                            meta=np.array(repr(dict(C=C, samples=samples, sigma=sigma, weight=weight, sizes=sizes,
                                                    seed=seed))))
                for i, s in enumerate(smp):
                    blob[f"sample.{i}"] = s.numpy()
            print(name, tag, "logged", logged, "|g_pred|max", float(g_pred.abs().max()), "|g_V|max", float(g_V.abs().max()))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)


if __name__ == "__main__":
    main()
