set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_transform.py -m gpu -x -q > gpurun_out/pytest_transform_r2b.log 2>&1; echo "transform tests exit $?"; tail -5 gpurun_out/pytest_transform_r2b.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or chunkwise or deterministic or decoder_backward_rows or out_of_range" > gpurun_out/pytest_new_r2b.log 2>&1; echo "new tests exit $?"; tail -5 gpurun_out/pytest_new_r2b.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu_r2b.log
timeout 300 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/smoke_r2b.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_r2b.log
timeout 900 python bench.py --no-partitioned > gpurun_out/bench_r2b.log 2>gpurun_out/bench_r2b.err; echo "bench exit $?"; tail -c 300 gpurun_out/bench_r2b.err
RGCN_PIPELINE=0 timeout 600 python bench.py --no-partitioned --quick --no-cpu-baseline > gpurun_out/bench_r2b_nopipe.log 2>&1; echo "bench nopipe $?"
PRIMEKG_RGCN_PREPARED_WEIGHTS=0 timeout 600 python bench.py --no-partitioned --quick --no-cpu-baseline > gpurun_out/bench_r2b_noprep.log 2>&1; echo "bench noprep $?"
timeout 600 python bench.py --no-partitioned --quick --no-cpu-baseline --mode bf16 > gpurun_out/bench_r2b_bf16.log 2>&1; echo "bench bf16 $?"
python - <<'PY'
import json
for f in ("bench_r2b","bench_r2b_nopipe","bench_r2b_noprep","bench_r2b_bf16"):
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1]);print(f,d["ms_per_step"],d["value"],d["e2e"]["value"], d['eager_ms_per_step'], d["dense_last_layer_bwd"]["ms_per_step"], d["gpu_launches_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-partitioned --quick > gpurun_out/ncu_list_r2b.log 2>&1; echo "ncu list $?"
timeout 300 python scripts/prof_timeline.py cfg2 > gpurun_out/timeline_cfg2_r2b.txt 2>&1; echo "timeline $?"
