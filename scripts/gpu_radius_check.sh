cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "radius_graph" 2>&1 | grep -E "radius_graph n=|passed|failed|Error|assert" | cut -c1-200
python - <<'PY'
import time, torch, numpy as np
from distegnn_b200 import radius_graph, synth
w = synth.WORKLOADS["synth1m"]
pts = synth.make_points(w, seed=0)
pos = torch.from_numpy(pts["pos"]).cuda()
for _ in range(2):
    ei, ea = radius_graph(pos, w.radius)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    ei, ea = radius_graph(pos, w.radius)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
t0 = time.perf_counter(); e_np = synth.radius_graph_np(pts["pos"], w.radius); t_cpu = time.perf_counter() - t0
print(f"RADIUS 1M nodes: device {dt*1e3:.2f} ms ({ei.shape[1]} edges) vs host cKDTree {t_cpu*1e3:.0f} ms ({e_np.shape[1]} edges)")
PY
