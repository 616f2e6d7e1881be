set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r2d.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_r2d.log
timeout 300 python scripts/ab_gemm.py > gpurun_out/ab_gemm_r2d.log 2>&1; echo "ab $?"; cat gpurun_out/ab_gemm_r2d.log
timeout 600 python bench.py --quick --no-cpu-baseline --no-partitioned > gpurun_out/bench_r2d.log 2>&1; echo "bench $?"
PRIMEKG_RGCN_PREPARED_WEIGHTS=0 timeout 600 python bench.py --quick --no-cpu-baseline --no-partitioned > gpurun_out/bench_r2d_noprep.log 2>&1; echo "bench $?"
python - <<'PY'
import json
for f in ("bench_r2d","bench_r2d_noprep"):
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1]);print(f,d["ms_per_step"],d["value"],d["e2e"]["value"], d['eager_ms_per_step'], d["dense_last_layer_bwd"]["ms_per_step"], d["gpu_launches_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
timeout 300 python scripts/prof_timeline.py cfg2 > gpurun_out/timeline_cfg2_r2d.txt 2>&1; echo "timeline $?"
sed -n 3,32p gpurun_out/timeline_cfg2_r2d.txt | cut -c1-100
