# Short multi-GPU check of the shipped build (gpurun --gpus N): the torchrun parity tests, then the N-GPU bench line
# (dist_parity inside, fused exchange + CUDA graph).  Usage: bash scripts/gpu_r02_multi_quick.sh <tag> <N>
cd $GRAFT_REPO_ROOT
TAG=${1:-r02q}
N=${2:-2}
timeout 600 python -m pytest tests -m gpu -q -s -k "torchrun or comm or fused_update" > gpurun_out/gpu_tests_multi_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/gpu_tests_multi_$TAG.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err; echo "benchN rc=$?"; tail -1 gpurun_out/bench_${N}gpu_$TAG.json | cut -c1-1500; tail -5 gpurun_out/bench_${N}gpu_$TAG.err | cut -c1-600
