# full GPU test suite on the default library, then one ncu --set full capture of the three per-layer kernels.
# Usage: bash scripts/gpu_tests_ncu.sh <tag>
cd $GRAFT_REPO_ROOT
TAG=${1:-t}
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/gpu_tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/gpu_tests_$TAG.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"edge_layer_cs|virtual_layer_t16|node_layer_tc" -s 3 -c 3 -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-train > gpurun_out/b_ncu_$TAG.log 2>&1; echo "ncu rc=$?"
