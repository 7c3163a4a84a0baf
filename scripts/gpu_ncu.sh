cd $GRAFT_REPO_ROOT
TAG=${1:-ncu}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"edge_layer_cs|virtual_layer_t16|node_layer_tc" -s 3 -c 3 -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-train > gpurun_out/b_ncu_$TAG.log 2>&1; echo "ncu rc=$?"
