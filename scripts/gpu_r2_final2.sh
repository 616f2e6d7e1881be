# round-2 final evidence (after the listed-rows last layer and the specialised transform epilogue): tests, smoke, bench lines
# (fp32 / bf16 / reference arm), ncu launch list of the bench command, ncu --set full of the listed walk and of the
# layer-1 forward transform, warm timeline
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final2.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu_final2.log
timeout 300 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/smoke_final2.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_final2.log
( time timeout 900 python bench.py > gpurun_out/bench_final2.log 2>gpurun_out/bench_final2.err ) 2>&1 | grep real; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_final2.log 2>&1; echo "ref arm $?"
timeout 600 python bench.py --mode bf16 --no-cpu-baseline --no-partitioned > gpurun_out/bench_bf16_final2.log 2>&1; echo "bf16 $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_final2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-partitioned --quick > gpurun_out/ncu_final2.log 2>&1; echo "ncu list $?"
timeout 300 ncu --set full --metrics lts__t_bytes.sum,lts__t_sectors_srcunit_tex.sum --clock-control none --import-source on -k regex:aggregate_rows_kernel --launch-skip 1 -c 1 -o gpurun_out/listed_walk_final python scripts/prof_listed.py > gpurun_out/ncu_listed_final.log 2>&1; echo "ncu listed $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_kmajor_kernel --launch-skip 1 -c 1 -o gpurun_out/gemm_l1_final python scripts/prof_gemm_l1.py > gpurun_out/ncu_gemm_l1_final.log 2>&1; echo "ncu gemm $?"
timeout 300 python scripts/prof_timeline.py cfg2 > gpurun_out/timeline_cfg2_final2.txt 2>&1
python - <<'PY'
import json
for f in ("bench_final2","bench_bf16_final2","bench_reference_final2"):
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1]); print(f, d.get("ms_per_step"), d.get("value"), (d.get("e2e") or {}).get("value"), d.get('eager_ms_per_step'), (d.get("dense_last_layer") or {}).get("ms_per_step"), (d.get("partitioned") or {}).get("ms_per_step"), d.get('cpu_baseline'))
    except Exception as e: print(f, "ERR", e)
PY
