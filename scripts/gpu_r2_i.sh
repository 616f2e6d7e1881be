set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_dist.py tests/test_gpu_rank.py -m gpu -q -x > gpurun_out/pytest_r2i.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_r2i.log
RGCN_PDL=0 timeout 600 python scripts/prof_partitioned.py > gpurun_out/prof_partitioned_r2i.txt 2>&1; echo "prof part $?"; tail -24 gpurun_out/prof_partitioned_r2i.txt
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r2i.log 2>&1; echo "bench $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2i.log").read().strip().splitlines()[-1]);print(d["ms_per_step"],d["value"],d["e2e"]["value"], d['eager_ms_per_step']); print(d.get("configs",{}).get("cfg4")); p=d["partitioned"]; print({k:p.get(k) for k in ("ms_per_step","nccl_exchange_ms_per_step","fused_vs_nccl_speedup","error","fused_error")})
PY
