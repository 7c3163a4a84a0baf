# Round-2 evidence run on an N-GPU box (gpurun --gpus N): GPU tests (incl. the 2-rank torchrun parity test), 1-GPU bench,
# N-GPU bench (dist_parity inside), N-GPU parity script logs.  Usage: bash scripts/gpu_r02_multi.sh <tag> <N>
set -x
cd $GRAFT_REPO_ROOT
TAG=${1:-r02}
N=${2:-2}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/gpu_tests_$TAG.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/gpu_tests_$TAG.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_1gpu_$TAG.json 2> gpurun_out/bench_1gpu_$TAG.err; echo "bench1 rc=$?"; cut -c1-600 gpurun_out/bench_1gpu_$TAG.json; tail -3 gpurun_out/bench_1gpu_$TAG.err
for mode in random kmeans; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/dist_parity.py --workload fluid113k --nodes 30000 --split-mode $mode --grads --cuda-graph > gpurun_out/dist_parity_${N}gpu_${mode}_$TAG.log 2>&1; echo "parity $mode rc=$?"; grep -E "^\{|PARITY" gpurun_out/dist_parity_${N}gpu_${mode}_$TAG.log | cut -c1-700
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err; echo "benchN rc=$?"; tail -1 gpurun_out/bench_${N}gpu_$TAG.json | cut -c1-1200; tail -5 gpurun_out/bench_${N}gpu_$TAG.err | cut -c1-600
DISTEGNN_B200_COMM=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-train --no-dist-parity > gpurun_out/bench_${N}gpu_nccl_$TAG.json 2> gpurun_out/bench_${N}gpu_nccl_$TAG.err; echo "benchN-nccl rc=$?"; tail -1 gpurun_out/bench_${N}gpu_nccl_$TAG.json | cut -c1-400
