set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r2c.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu_r2c.log
timeout 300 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/smoke_r2c.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_r2c.log
timeout 900 python bench.py > gpurun_out/bench_r2c.log 2>gpurun_out/bench_r2c.err; echo "bench exit $?"; tail -c 300 gpurun_out/bench_r2c.err
RGCN_PIPELINE=0 timeout 600 python bench.py --quick --no-cpu-baseline > gpurun_out/bench_r2c_nopipe.log 2>&1; echo "bench nopipe $?"
timeout 600 python bench.py --no-partitioned --quick --no-cpu-baseline --mode bf16 > gpurun_out/bench_r2c_bf16.log 2>&1; echo "bench bf16 $?"
python - <<'PY'
import json
for f in ("bench_r2c","bench_r2c_nopipe","bench_r2c_bf16"):
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1]);print(f,d["ms_per_step"],d["value"],d["e2e"]["value"], d['eager_ms_per_step'], d["dense_last_layer_bwd"]["ms_per_step"], d["gpu_launches_per_step"]); p=d.get("partitioned") or {}; print("  part", p.get("ms_per_step"), p.get("nccl_exchange_ms_per_step"), p.get("error"))
    except Exception as e: print(f, "ERR", e)
PY
timeout 300 python scripts/prof_timeline.py cfg2 > gpurun_out/timeline_cfg2_r2c.txt 2>&1; echo "timeline $?"
