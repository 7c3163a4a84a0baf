"""Shared helpers for the test-suite (fixtures loader, error metrics)."""
from __future__ import annotations

import ast
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SINGLE_CASES = ["nbody24_norm", "fluid160_c5", "batch3_c8_multigraph"]
DIST_CASE = "dist2_fluid300_c5"
INPUT_KEYS = ["node_feat", "node_loc", "node_vel", "loc_mean", "edge_index", "data_batch", "edge_attr",
              "node_attr"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    kw = ast.literal_eval(str(z["meta.kw"]))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    return z, kw, sd


def golden_inputs(z, prefix="in."):
    d = {k: (torch.from_numpy(z[prefix + k]) if prefix + k in z.files else None) for k in INPUT_KEYS}
    return d


def golden_trace(z, key, prefix="trace."):
    out, i = [], 0
    while f"{prefix}{key}.{i}" in z.files:
        out.append(torch.from_numpy(z[f"{prefix}{key}.{i}"]))
        i += 1
    return out


def max_abs(a, b):
    return float((a.double() - b.double()).abs().max()) if a.numel() else 0.0


def rel_disp_err(out, ref, pos):
    """‖(out−pos)−(ref−pos)‖∞ / ‖ref−pos‖∞ — parity on the *displacement*, which is what the model
    actually computes (SURVEY §7 'parity is deceptively easy at init')."""
    den = float((ref.double() - pos.double()).abs().max())
    return max_abs(out, ref) / max(den, 1e-30)
