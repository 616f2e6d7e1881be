set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_transform.py tests/test_gpu_rank.py -m gpu -q -x > gpurun_out/pytest_r2e_a.log 2>&1; echo "transform+rank exit $?"; tail -15 gpurun_out/pytest_r2e_a.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r2e.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu_r2e.log
timeout 300 python scripts/ab_gemm.py > gpurun_out/ab_gemm_r2e.log 2>&1; echo "ab $?"; cat gpurun_out/ab_gemm_r2e.log
timeout 600 python bench.py --no-cpu-baseline --no-partitioned > gpurun_out/bench_r2e.log 2>&1; echo "bench $?"
RGCN_WIDE_TILES=0 timeout 600 python bench.py --quick --no-cpu-baseline --no-partitioned > gpurun_out/bench_r2e_narrow.log 2>&1; echo "bench $?"
python - <<'PY'
import json
for f in ("bench_r2e","bench_r2e_narrow"):
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1]);print(f,d["ms_per_step"],d["value"],d["e2e"]["value"], d['eager_ms_per_step'], d["dense_last_layer_bwd"]["ms_per_step"], d["gpu_launches_per_step"]); print(d.get("configs",{}).get("cfg4")); print(d["roofline"]["tensor"]["modes"])
    except Exception as e: print(f, "ERR", e)
PY
