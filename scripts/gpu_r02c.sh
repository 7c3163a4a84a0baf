cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_loss.py -m gpu -x -q -s 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "radius_graph_csr or device_built or kmeans or split_large or edge_stage_backward or golden or shard" 2>&1 | tail -25
bash scripts/gpu_variants_bench.sh r02c
timeout 200 python main.py --config_path config/largefluid_distegnn.yaml --eval_steps 3 --train_steps 4 2>&1 | tail -6
