set -x
mkdir -p gpurun_out
timeout 600 python scripts/diag_hub_bwd.py > gpurun_out/diag_hub_bwd.log 2>&1; echo "diag $?"; cat gpurun_out/diag_hub_bwd.log
timeout 900 ncu --set full --clock-control none -k regex:"hub_partial" -s 2 -c 2 -f -o gpurun_out/prof_hub_bwd_r2 python scripts/diag_hub_bwd.py > gpurun_out/prof_hub_bwd_ncu.log 2>&1; echo "ncu $?"
