cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/profile_train_step.py 2>&1 | grep -E "Name|bwd|FastEGNNFunctionBackward  |Self CUDA time total|aten::mm "
