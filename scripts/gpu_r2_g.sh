set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r2g.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_r2g.log
timeout 600 python bench.py --no-cpu-baseline --no-partitioned > gpurun_out/bench_r2g.log 2>&1; echo "bench $?"
python - <<'PY'
import json
for f in ("bench_r2g",):
    d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1]);print(f,d["ms_per_step"],d["value"],d["e2e"]["value"], d['eager_ms_per_step'], d["dense_last_layer_bwd"]["ms_per_step"], d["gpu_launches_per_step"]); print(d.get("configs",{}).get("cfg4")); print(d.get("configs",{}).get("cfg3")); print(d.get("configs",{}).get("cfg1"))
PY
timeout 600 python scripts/prof_partitioned.py > gpurun_out/prof_partitioned_r2g.txt 2>&1; echo "prof part $?"; tail -30 gpurun_out/prof_partitioned_r2g.txt
timeout 300 python scripts/prof_timeline.py cfg2 > gpurun_out/timeline_cfg2_r2g.txt 2>&1; sed -n 3,32p gpurun_out/timeline_cfg2_r2g.txt | cut -c1-100
