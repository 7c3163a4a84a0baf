cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/gpu_tests_r02c.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/gpu_tests_r02c.log | cut -c1-300
bash scripts/gpu_variants_bench.sh r02c
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_1gpu_r02c.json 2> gpurun_out/bench_1gpu_r02c.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_1gpu_r02c.json; tail -3 gpurun_out/bench_1gpu_r02c.err
timeout 200 python main.py --config_path config/largefluid_distegnn.yaml --eval_steps 3 --train_steps 4 2>&1 | tail -6
