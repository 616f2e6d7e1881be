cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused_layer.py -x -q -m gpu -k "aggregate or micro or csr or fused_layer or golden" 2>&1 | tail -4
timeout 300 python scripts/bench_agg.py 2>&1 | tail -12
timeout 600 python bench.py --steps 20 --warmup 3 --no-partitioned --no-cpu-baseline > gpurun_out/bench_div.log 2>&1; echo "bench rc $?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_div.log'):
    if l.startswith('{'):
        d=json.loads(l); print('ms', d['ms_per_step'], 'eager', d['eager_ms_per_step'], 'roof', d['roofline']['avg_launch_ms'], d['roofline']['frac'], 'd64', d['roofline']['d64'])
PY
