# A/B harness for one gpurun call: quick GPU tests on the default library (under a short timeout, so that a hung kernel
# costs minutes and not the whole call), then kernel timings for the default library and every production variant under
# distegnn_b200/variants/ (python -m distegnn_b200.build --variant TAG --defs=...), optionally one ncu capture.
# Usage: bash scripts/gpu_ab_run.sh <tag> "<pytest -k expr>" [ncu]
set -x
cd $GRAFT_REPO_ROOT
TAG=${1:-ab}
KEXPR=${2:-"edge_kernel or virtual_kernel or library_loaded or gather4 or golden or workloads"}
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "$KEXPR" > gpurun_out/gpu_tests_quick_$TAG.log 2>&1
rc=$?; echo "quick tests rc=$rc"; tail -12 gpurun_out/gpu_tests_quick_$TAG.log
for lib in default $(ls distegnn_b200/variants/libdistegnn_b200.*.so 2>/dev/null | grep -v _testing); do
  if [ $lib = default ]; then unset DISTEGNN_B200_LIB; else export DISTEGNN_B200_LIB=$GRAFT_REPO_ROOT/$lib; fi
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-train 2>gpurun_out/ab_err_$TAG.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('VARIANT $lib', round(d['ms_per_step'],3), {k:round(v,4) for k,v in d['kernel_ms'].items() if k!='note'}, 'roofline', round(d['roofline']['frac'],3))" || tail -5 gpurun_out/ab_err_$TAG.log
done
unset DISTEGNN_B200_LIB
if [ "$3" = "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"edge_layer_cs|virtual_layer_t16" -s 2 -c 2 -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-train > gpurun_out/b_ncu_$TAG.log 2>&1; echo ncu rc=$?
fi
