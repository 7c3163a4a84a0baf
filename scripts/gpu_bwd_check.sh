cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stage_backward or training" 2>&1 | tail -2
timeout 300 python scripts/profile_train_step.py 2>&1 | grep -E "bwd|Self CUDA time total"
