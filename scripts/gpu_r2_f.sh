set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r2f.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu_r2f.log
timeout 600 python bench.py --no-cpu-baseline --no-partitioned --mode bf16 > gpurun_out/bench_r2f_bf16.log 2>&1; echo "bench $?"
PRIMEKG_RGCN_BF16_GATHER=0 timeout 600 python bench.py --quick --no-cpu-baseline --no-partitioned --mode bf16 > gpurun_out/bench_r2f_bf16_f32gather.log 2>&1; echo "bench $?"
timeout 600 python bench.py --quick --no-cpu-baseline --no-partitioned > gpurun_out/bench_r2f.log 2>&1; echo "bench $?"
python - <<'PY'
import json
for f in ("bench_r2f_bf16","bench_r2f_bf16_f32gather","bench_r2f"):
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1]);print(f,d["ms_per_step"],d["value"],d["e2e"]["value"], d['eager_ms_per_step'], d["dense_last_layer_bwd"]["ms_per_step"], d["gpu_launches_per_step"]); print(d.get("configs",{}).get("cfg4")); print(d.get("configs",{}).get("cfg3"))
    except Exception as e: print(f, "ERR", e)
PY
timeout 300 python scripts/prof_eager_host.py > gpurun_out/eager_host_r2f.log 2>&1; echo "eager prof $?"; head -60 gpurun_out/eager_host_r2f.log
