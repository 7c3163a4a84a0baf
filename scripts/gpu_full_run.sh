# Full evidence run for one gpurun call: GPU tests, the default bench line, the ncu launch list of the same command,
# one `ncu --set full` capture of the dominant kernels, and the torchrun-style entry point with training steps.
# Usage: bash scripts/gpu_full_run.sh <tag>
set -x
cd $GRAFT_REPO_ROOT
TAG=${1:-full}
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/gpu_tests_$TAG.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/gpu_tests_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json | cut -c1-1500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-train > gpurun_out/b_launch_$TAG.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"edge_layer_cs|virtual_layer_t16|node_layer_tc" -s 3 -c 3 -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-train > gpurun_out/b_ncu_$TAG.log 2>&1; echo "ncu rc=$?"
timeout 300 python main.py --config_path config/largefluid_distegnn.yaml --eval_steps 5 --train_steps 6 2>&1 | tail -12
