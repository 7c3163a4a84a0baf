cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "shard_input" 2>&1 | grep -E "shard:|passed|failed|Error|assert" | cut -c1-200
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-train 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); e=d['e2e']; print('E2E', round(e['ms_per_step'],2), {k:(round(e[k]['ms_per_step'],2), e[k].get('h2d_bytes_per_step')) for k in ('from_shard','from_positions','pipelined')})"
