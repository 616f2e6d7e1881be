timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
for o in 0 1; do RGCN_OVERLAP_WGRAD=$o python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ow$o.log 2>&1; echo "bench overlap=$o exit $?"; done
python - <<'PY'
import json
for f in ("gpurun_out/bench_ow0.log","gpurun_out/bench_ow1.log"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]);print(f,d["ms_per_step"],d["value"],d["e2e"]["value"], d['eager_ms_per_step'], d['gpu_launches_per_step'], d['dense_last_layer_bwd'])
    except Exception as e:
        print(f, 'ERR', e); print(open(f).read()[-2000:])
PY
