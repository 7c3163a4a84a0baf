# kernel timings of the default library and every variant at config-5 size AND at 125k nodes (the per-GPU share at N=8),
# CUDA graph on.  Usage: bash scripts/gpu_small_bench.sh <tag>
cd $GRAFT_REPO_ROOT
TAG=${1:-s}
for lib in default $(ls distegnn_b200/variants/libdistegnn_b200.*.so 2>/dev/null | grep -v _testing); do
  if [ $lib = default ]; then unset DISTEGNN_B200_LIB; else export DISTEGNN_B200_LIB=$GRAFT_REPO_ROOT/$lib; fi
  for nodes in 1000000 125000; do
  timeout 300 python bench.py --nodes $nodes --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-train 2>gpurun_out/small_err_$TAG.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('VARIANT $lib nodes=$nodes', round(d['ms_per_step'],3), {k:round(v,4) for k,v in d['kernel_ms'].items() if k!='note'})" >> gpurun_out/small_$TAG.txt 2>&1 || tail -5 gpurun_out/small_err_$TAG.log >> gpurun_out/small_$TAG.txt
  done
done
cat gpurun_out/small_$TAG.txt
