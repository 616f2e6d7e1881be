# round-2 final evidence: tests, smoke, bench lines (fp32 / bf16 / reference arm), ncu launch list of the bench command, warm timeline
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_final.log
( time timeout 900 python bench.py > gpurun_out/bench_final.log 2>gpurun_out/bench_final.err ) 2>&1 | grep real; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_final.log 2>&1; echo "ref arm $?"
timeout 600 python bench.py --mode bf16 --no-cpu-baseline --no-partitioned > gpurun_out/bench_bf16_final.log 2>&1; echo "bf16 $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-partitioned > gpurun_out/ncu_final.log 2>&1; echo "ncu list $?"
timeout 300 python scripts/prof_timeline.py cfg2 > gpurun_out/timeline_cfg2_final.txt 2>&1
python - <<'PY'
import json
for f in ("bench_final","bench_bf16_final","bench_reference_final"):
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1]); print(f, d.get("ms_per_step"), d.get("value"), (d.get("e2e") or {}).get("value"), d.get('eager_ms_per_step'), (d.get("dense_last_layer_bwd") or {}).get("ms_per_step"), d.get('cpu_baseline'))
    except Exception as e: print(f, "ERR", e)
PY
