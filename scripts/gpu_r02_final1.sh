# One-call evidence run on 1 GPU (round 2): A/B of the edge-kernel knobs, then everything else on the FASTEST build:
# full GPU tests, default bench line, ncu launch list + full capture, full-size parity vs the reference's CPU forward,
# the reference arm of the bench (skipped with SKIP_REF=1), smoke().  Usage: bash scripts/gpu_r02_final1.sh <tag>
cd $GRAFT_REPO_ROOT
TAG=${1:-r02f}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv
rm -f gpurun_out/variants_$TAG.txt
bash scripts/gpu_variants_bench.sh $TAG
BEST=$(python - <<PY
import re,ast
best=None
for l in open("gpurun_out/variants_$TAG.txt"):
    m=re.match(r"VARIANT (\S+) ([\d.]+) (\{.*?\}) roofline", l)
    if m:
        e=ast.literal_eval(m.group(3))["edge"]
        if best is None or e<best[0]-0.01: best=(e,m.group(1))
print(best[1] if best else "default")
PY
)
echo "BEST_VARIANT $BEST" | tee -a gpurun_out/variants_$TAG.txt
if [ "$BEST" != "default" ]; then export DISTEGNN_B200_LIB=$GRAFT_REPO_ROOT/$BEST; fi
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/gpu_tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/gpu_tests_$TAG.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench_1gpu_$TAG.json 2> gpurun_out/bench_1gpu_$TAG.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_1gpu_$TAG.json; tail -2 gpurun_out/bench_1gpu_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-train > gpurun_out/b_launch_$TAG.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"edge_layer_cs|virtual_layer_t16|node_layer_tc" -s 3 -c 3 -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-train > gpurun_out/b_ncu_$TAG.log 2>&1; echo "ncu rc=$?"
timeout 400 python scripts/full_size_oracle_check.py > gpurun_out/full_size_parity_$TAG.json 2> gpurun_out/full_size_parity_$TAG.err; echo "full-size rc=$?"; cat gpurun_out/full_size_parity_$TAG.json
[ -n "$SKIP_REF" ] || timeout 700 python bench.py --impl reference --steps 20 --warmup 5 --ref-budget-s 200 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; cut -c1-700 gpurun_out/bench_ref_$TAG.json; tail -2 gpurun_out/bench_ref_$TAG.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 200 python main.py --config_path config/largefluid_distegnn.yaml --eval_steps 3 --train_steps 4 2>&1 | tail -4
