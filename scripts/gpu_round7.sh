timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rows or sparse or graphed or golden" > gpurun_out/pytest_sparse.log 2>&1; echo "pytest-sparse exit $?"; tail -15 gpurun_out/pytest_sparse.log
for s in 0 1; do PRIMEKG_RGCN_SPARSE_BWD=$s python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sparse$s.log 2>&1; echo "bench sparse=$s exit $?"; done
python - <<'PY'
import json
for f in ("gpurun_out/bench_sparse0.log","gpurun_out/bench_sparse1.log"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]);print(f,d["ms_per_step"],d["value"],d["e2e"]["value"], d['eager_ms_per_step'], d['gpu_launches_per_step'])
    except Exception as e:
        print(f, 'ERR', e); print(open(f).read()[-2000:])
PY
