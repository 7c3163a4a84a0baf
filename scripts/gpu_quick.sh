cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "virtual_kernel or golden or edge_cases or batched" 2>&1 | tail -8
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('BENCH', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms'].items()})"
