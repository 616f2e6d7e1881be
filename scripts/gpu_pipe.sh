timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "aggregate or micro or rows or planes or hub or full_size or sparse" > gpurun_out/pytest_pipe.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_pipe.log
for r in 0 2 4 8; do echo "PIPE=$r"; RGCN_AGG_PIPE=$r python scripts/bench_agg.py 2>&1 | grep -v hubs | grep -v "fp32-out"; done
for r in 0 2 4; do RGCN_AGG_PIPE=$r python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pipe$r.log 2>&1; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_pipe$r.log").read().strip().splitlines()[-1]);print("PIPE=$r",d["ms_per_step"],d["value"],d["dense_last_layer_bwd"]["ms_per_step"])
PY
done
