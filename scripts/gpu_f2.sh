timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu.log
python scripts/bench_agg.py 2>&1 | grep -v hubs | grep -v "fp32-out"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32.log 2>&1; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_fp32.log").read().strip().splitlines()[-1]);print(d["ms_per_step"],d["value"],d["e2e"]["value"],d["dense_last_layer_bwd"]["ms_per_step"])
PY
python scripts/bench_cfg.py cfg1 | cut -c1-200; python scripts/bench_cfg.py cfg3 | cut -c1-230
