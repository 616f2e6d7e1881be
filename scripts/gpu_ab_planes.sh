for h in 0 1; do PRIMEKG_RGCN_PLANES_HANDOVER=$h python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ph$h.log 2>&1; echo "bench handover=$h exit $?"; done
python - <<'PY'
import json
for f in ("gpurun_out/bench_ph0.log","gpurun_out/bench_ph1.log"):
    d=json.loads(open(f).read().strip().splitlines()[-1]);print(f,d["ms_per_step"],d["value"],d["e2e"]["value"], d['eager_ms_per_step'], d["gpu_launches_per_step"], d["dense_last_layer_bwd"]["ms_per_step"])
PY
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
timeout 300 python scripts/prof_timeline.py cfg2 > gpurun_out/timeline_cfg2_sparse.txt 2>&1; grep -n "aggregate_rows\|hub_partial\|split_planes" gpurun_out/timeline_cfg2_sparse.txt | tail -9 | cut -c1-120
