timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
for v in 1 0; do
RGCN_PDL=$v python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl$v.log 2>&1; echo "bench pdl=$v exit $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_pdl$v.log").read().strip().splitlines()[-1]);print("pdl=$v",d["ms_per_step"],d["value"],d["e2e"]["value"], d["e2e"]["eager_module_api_value"], d["eager_ms_per_step"])
PY
done
for v in 1 0; do RGCN_PDL=$v python scripts/bench_cfg.py cfg1 | cut -c1-200; RGCN_PDL=$v python scripts/bench_cfg.py cfg3 | cut -c1-230; done
