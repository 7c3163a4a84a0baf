# A/B harness for one gpurun call: GPU tests on the default library (edge-kernel tests first, under a short timeout, so
# that a hung kernel costs two minutes and not the whole call), then kernel timings for every variant library under
# distegnn_b200/variants/, then one ncu capture of the two dominant kernels.  Usage: bash scripts/gpu_ab_run.sh [tag]
set -x
cd $GRAFT_REPO_ROOT
TAG=${1:-ab}
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "edge_kernel or virtual_kernel or library_loaded" > gpurun_out/gpu_tests_quick_$TAG.log 2>&1
rc=$?; echo "quick tests rc=$rc"; tail -15 gpurun_out/gpu_tests_quick_$TAG.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/gpu_tests_$TAG.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/gpu_tests_$TAG.log
for lib in default distegnn_b200/variants/*.so; do
  if [ $lib = default ]; then unset DISTEGNN_B200_LIB; else export DISTEGNN_B200_LIB=$GRAFT_REPO_ROOT/$lib; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-train 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('VARIANT $lib', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms'].items()})"
done
unset DISTEGNN_B200_LIB
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"edge_layer_cs|virtual_layer_t16" -s 2 -c 2 -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-train > gpurun_out/b_ncu_$TAG.log 2>&1; echo ncu rc=$?
