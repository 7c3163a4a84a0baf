# Multi-GPU evidence on one N-GPU box (gpurun --gpus N): the 2-rank torchrun parity tests, dist_parity at N ranks (random +
# k-means, graph, gradients), bench at every power of two <= N with the fused exchange (+ the NCCL path at N), DDP training
# through main.py at N ranks.  Usage: bash scripts/gpu_r02_scale.sh <tag> <N>
cd $GRAFT_REPO_ROOT
TAG=${1:-r02s}
N=${2:-8}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "multi_gpu" > gpurun_out/gpu_tests_multigpu_$TAG.log 2>&1; echo "torchrun tests rc=$?"; tail -3 gpurun_out/gpu_tests_multigpu_$TAG.log | cut -c1-300
for mode in random kmeans; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/dist_parity.py --workload fluid113k --nodes 40000 --split-mode $mode --grads --cuda-graph > gpurun_out/dist_parity_${N}gpu_${mode}_$TAG.log 2>&1; echo "parity $mode rc=$?"; grep -E "^\{|PARITY" gpurun_out/dist_parity_${N}gpu_${mode}_$TAG.log | cut -c1-700
done
n=$N
while [ $n -ge 2 ]; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_${n}gpu_$TAG.json 2> gpurun_out/bench_${n}gpu_$TAG.err; echo "bench N=$n rc=$?"; tail -1 gpurun_out/bench_${n}gpu_$TAG.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value'],1), {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['kernel_ms'].items() if k!='note'}, 'graph', d['config']['cuda_graph'], 'parity', d['dist_parity']['pass'], 'train', round(d['train_step']['ms_per_step'],1), 'e2e', round(d['e2e']['ms_per_step'],2), 'sum_edges', d['config']['edges_total_sum_p'])"; tail -2 gpurun_out/bench_${n}gpu_$TAG.err | cut -c1-300
n=$((n/2))
done
DISTEGNN_B200_COMM=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-train --no-dist-parity > gpurun_out/bench_${N}gpu_nccl_$TAG.json 2> gpurun_out/bench_${N}gpu_nccl_$TAG.err; echo "bench nccl rc=$?"; tail -1 gpurun_out/bench_${N}gpu_nccl_$TAG.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('NCCL N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['kernel_ms'].items() if k!='note'})"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29536 main.py --config_path config/largefluid_distegnn.yaml --eval_steps 5 --train_steps 6 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -8
