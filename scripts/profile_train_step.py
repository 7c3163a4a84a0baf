"""Where a train step (forward + backward) spends its GPU time: torch.profiler table of one step.
    python scripts/profile_train_step.py [--workload synth1m] [--nodes N]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distegnn_b200 import FastEGNN, synth  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="synth1m")
ap.add_argument("--nodes", type=int, default=None)
a = ap.parse_args()
w = synth.WORKLOADS[a.workload]
dev = torch.device("cuda", 0)
host = synth.make_partitions(w, n_nodes=a.nodes or w.n_nodes, seed=0)[0]
inp = {k: (v.to(dev) if v is not None else None) for k, v in host.items()}
model = FastEGNN(hidden_nf=64, world_size=1, **bench.model_dims(w))
model.load_state_dict(bench.make_state_dict(w))
model = model.to(dev).train()
target = inp["node_loc"] + 0.01 * inp["node_vel"]


def step():
    for p in model.parameters():
        p.grad = None
    out, X = model(**inp)
    (torch.nn.functional.mse_loss(out, target) + 1e-3 * X.square().mean()).backward()


step(); step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
