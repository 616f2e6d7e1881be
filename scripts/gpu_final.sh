timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu.log
RGCN_OVERLAP_WGRAD=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_transform.py -m gpu -x -q -k "golden or graphed or sparse or dropout or full_size or transform or layer" > gpurun_out/pytest_forced_overlap.log 2>&1; echo "pytest (forced side stream) exit $?"; tail -2 gpurun_out/pytest_forced_overlap.log
python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32.log 2>&1; echo "bench exit $?"
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --mode bf16 > gpurun_out/bench_bf16.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/bench_fp32.log","gpurun_out/bench_bf16.log"):
    d=json.loads(open(f).read().strip().splitlines()[-1]);print(f,d["ms_per_step"],d["value"],d["e2e"]["value"], d['eager_ms_per_step'], d["gpu_launches_per_step"], d["dense_last_layer_bwd"]["ms_per_step"])
PY
timeout 300 python scripts/prof_timeline.py cfg2 > gpurun_out/timeline_cfg2.txt 2>&1
