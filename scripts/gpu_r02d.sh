cd $GRAFT_REPO_ROOT
bash scripts/gpu_variants_bench.sh r02d
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "edge_kernel or golden or workloads or fused_sync or validation or accumulators or edge_cases or gather4 or equivariance" 2>&1 | tail -3
timeout 400 python scripts/full_size_oracle_check.py > gpurun_out/full_size_parity_r02d.json 2> gpurun_out/full_size_parity_r02d.err; echo "full-size rc=$?"; cat gpurun_out/full_size_parity_r02d.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 --ref-budget-s 150 > gpurun_out/bench_ref_r02d.json 2> gpurun_out/bench_ref_r02d.err; echo "ref rc=$?"; cut -c1-900 gpurun_out/bench_ref_r02d.json; tail -3 gpurun_out/bench_ref_r02d.err
