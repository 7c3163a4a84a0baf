cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/dist_parity.py --workload fluid113k --nodes 20000 --grads 2>&1 | grep -E "rank|PARITY|gradients|Error|error" | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_2gpu_r01k.json; cut -c1-900 gpurun_out/bench_2gpu_r01k.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 main.py --config_path config/largefluid_distegnn.yaml --eval_steps 5 --train_steps 5 2>&1 | tail -8
