timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32.log 2>&1; echo "bench exit $?"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --mode bf16 > gpurun_out/bench_bf16.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/bench_fp32.log","gpurun_out/bench_bf16.log"):
    d=json.loads(open(f).read().strip().splitlines()[-1]);print(f,d["ms_per_step"],d["value"],d["e2e"]["value"], d['eager_ms_per_step'])
PY
python scripts/bench_cfg.py cfg1 | cut -c1-200; python scripts/bench_cfg.py cfg3 | cut -c1-230
python scripts/diag_transform.py 2>&1 | grep "cfg2.*fp32" | tail -8
