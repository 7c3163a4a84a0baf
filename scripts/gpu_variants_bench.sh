# kernel timings of the default library and of every production variant under distegnn_b200/variants/ (no tests, no ncu);
# results in gpurun_out/variants_<tag>.txt.  Usage: bash scripts/gpu_variants_bench.sh <tag>
cd $GRAFT_REPO_ROOT
TAG=${1:-v}
for lib in default $(ls distegnn_b200/variants/libdistegnn_b200.*.so 2>/dev/null | grep -v _testing); do
  if [ $lib = default ]; then unset DISTEGNN_B200_LIB; else export DISTEGNN_B200_LIB=$GRAFT_REPO_ROOT/$lib; fi
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-train 2>gpurun_out/variants_err_$TAG.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('VARIANT $lib', round(d['ms_per_step'],3), {k:round(v,4) for k,v in d['kernel_ms'].items() if k!='note'}, 'roofline', round(d['roofline']['frac'],3))" >> gpurun_out/variants_$TAG.txt 2>&1 || tail -5 gpurun_out/variants_err_$TAG.log >> gpurun_out/variants_$TAG.txt
done
cat gpurun_out/variants_$TAG.txt
