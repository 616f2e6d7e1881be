python scripts/bench_agg.py > gpurun_out/bench_agg.log 2>&1; echo "agg plain $?"; cat gpurun_out/bench_agg.log
ncu --set full --clock-control none --import-source on -k regex:"aggregate_rows_kernel|hub_partial" -s 20 -c 10 -f -o gpurun_out/prof_agg_r1c python scripts/bench_agg.py > gpurun_out/prof_agg_ncu.log 2>&1; echo "ncu agg $?"
python scripts/prof_gemm.py > gpurun_out/prof_gemm_plain.log 2>&1; echo "gemm plain $?"
ncu --set full --clock-control none --import-source on -k regex:"gemm_" -s 6 -c 3 -f -o gpurun_out/prof_gemm_r1c python scripts/prof_gemm.py > gpurun_out/prof_gemm_ncu.log 2>&1; echo "ncu gemm $?"
