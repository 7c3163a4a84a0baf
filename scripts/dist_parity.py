"""Multi-GPU parity: run under torchrun with N ranks (one per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        scripts/dist_parity.py [--workload fluid113k] [--nodes 113140] [--split-mode random|kmeans] [--grads] [--cuda-graph]

Thin command-line wrapper of oracle/dist_check.py (the checker bench.py also runs before its timed region): every rank
runs its partition through the CUDA path with the real cross-rank exchange, rank 0 compares all ranks' outputs with the
partitioned float64 oracle and prints one JSON line + `DIST_PARITY PASS|FAIL`.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="fluid113k")
    ap.add_argument("--nodes", type=int, default=None)
    ap.add_argument("--split-mode", default="random")
    ap.add_argument("--grads", action="store_true", help="also check the training path (use a small --nodes: the "
                    "oracle's float64 autograd runs on the host)")
    ap.add_argument("--cuda-graph", action="store_true", help="run the forward as a captured CUDA graph (replayed once)")
    args = ap.parse_args()
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from oracle import dist_check
    res = dist_check.check_case(args.workload, args.nodes, args.split_mode, dev, grads=args.grads,
                                cuda_graph=args.cuda_graph)
    if rank == 0:
        print(json.dumps(res), flush=True)
        print("DIST_PARITY", "PASS" if res["pass"] else "FAIL",
              f"world={res['world']} split={args.split_mode} {args.workload}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not res["pass"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
