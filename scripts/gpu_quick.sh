# quick GPU check: bash scripts/gpu_quick.sh "<pytest -k expression>"
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "$1" 2>&1 | tail -25
