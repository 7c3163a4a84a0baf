cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "stage_backward or training" 2>&1 | grep -E "virtual stage|gradients|passed|failed|Error" | cut -c1-260
timeout 300 python scripts/profile_train_step.py 2>&1 | grep -E "Name|bwd_kernel|FastEGNNFunctionBackward  |Self CUDA time total|aten::mm "
