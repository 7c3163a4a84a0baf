# quick GPU check: bash scripts/gpu_quick.sh "<pytest -k expression>" [extra command]
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "$1" 2>&1 | tail -25
if [ -n "$2" ]; then bash -c "$2"; fi
