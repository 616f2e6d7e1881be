set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "ref arm $?"
python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench $?"
python bench.py --mode bf16 --no-cpu-baseline > gpurun_out/bench_bf16.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/bench_default.log","gpurun_out/bench_bf16.log"):
    d=json.loads(open(f).read().strip().splitlines()[-1]);print(f,d["ms_per_step"],d["value"],d["e2e"]["value"], d['eager_ms_per_step'], d["dense_last_layer_bwd"]["ms_per_step"], d.get('cpu_baseline'))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo "ncu list $?"
python scripts/prof_agg256.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"aggregate_rows_kernel|hub_partial" -s 8 -c 4 -f -o gpurun_out/prof_agg256_r1 python scripts/prof_agg256.py > gpurun_out/prof_agg_ncu.log 2>&1; echo "ncu agg $?"
python scripts/prof_rowsparse.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"aggregate_rows_kernel|hub_partial|rows_|gemm_" -s 30 -c 18 -f -o gpurun_out/prof_rowsparse_r1 python scripts/prof_rowsparse.py > gpurun_out/prof_rowsparse_ncu.log 2>&1; echo "ncu rowsparse $?"
timeout 300 python scripts/prof_timeline.py cfg2 > gpurun_out/timeline_cfg2.txt 2>&1
for c in cfg1 cfg3; do python scripts/bench_cfg.py $c | cut -c1-260; done
python scripts/bench_cfg.py cfg3 bf16 | cut -c1-260
