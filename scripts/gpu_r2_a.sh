set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv
lscpu | grep -i "model name\|^CPU(s)" 
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu_r2a.log
timeout 300 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/smoke_r2a.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_r2a.log
timeout 900 python bench.py > gpurun_out/bench_r2a.log 2>gpurun_out/bench_r2a.err; echo "bench exit $?"; tail -c 600 gpurun_out/bench_r2a.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2a.log 2>&1; echo "ref arm $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r2a.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-partitioned --quick > gpurun_out/ncu_list_r2a.log 2>&1; echo "ncu list $?"
timeout 600 ncu --set full --metrics lts__t_bytes.sum,lts__t_sectors_srcunit_tex.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none --import-source on -k regex:"aggregate_rows_kernel|hub_partial" -s 8 -c 4 -f -o gpurun_out/prof_agg256_r2 python scripts/prof_agg256.py > gpurun_out/prof_agg_ncu_r2.log 2>&1; echo "ncu agg $?"
timeout 600 ncu --set full --metrics lts__t_bytes.sum,lts__t_sectors_srcunit_tex.sum --clock-control none -k regex:"aggregate_rows_kernel|probe_gather" -s 2 -c 2 -f -o gpurun_out/prof_agg_hbm_r2 python scripts/prof_agg_hbm.py > gpurun_out/prof_agg_hbm_ncu_r2.log 2>&1; echo "ncu hbm $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_" -s 6 -c 3 -f -o gpurun_out/prof_gemm_r2 python scripts/prof_gemm.py > gpurun_out/prof_gemm_ncu_r2.log 2>&1; echo "ncu gemm $?"
ls -la gpurun_out/*.ncu-rep | tail -5
