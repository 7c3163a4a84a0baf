cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 scripts/dist_parity.py --workload fluid113k --nodes 40000 --split-mode kmeans --grads 2>&1 | grep -E "rank|PARITY|gradients|Error|error" | tail -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_8gpu_r01n.json; cut -c1-700 gpurun_out/bench_8gpu_r01n.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 main.py --config_path config/synth1m_distegnn.yaml --eval_steps 5 --train_steps 5 2>&1 | grep -vE "Warning|warn|reducer|^$" | tail -8
