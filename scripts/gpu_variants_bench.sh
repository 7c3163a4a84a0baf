# kernel timings of the default library and of every variant under distegnn_b200/variants/ (no tests, no ncu)
cd $GRAFT_REPO_ROOT
for lib in default distegnn_b200/variants/*.so; do
  if [ $lib = default ]; then unset DISTEGNN_B200_LIB; else export DISTEGNN_B200_LIB=$GRAFT_REPO_ROOT/$lib; fi
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('VARIANT $lib', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms'].items()}, 'train', round(d['train_step']['ms_per_step'],1))"
done
