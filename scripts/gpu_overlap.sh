timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
for v in 1 0; do
echo "== RGCN_OVERLAP_HUBS=$v"
RGCN_OVERLAP_HUBS=$v python scripts/bench_agg.py 2>&1 | grep -v "fp32-out\|hubs fwd"
RGCN_OVERLAP_HUBS=$v python scripts/bench_cfg.py cfg2 | cut -c1-200
RGCN_OVERLAP_HUBS=$v python scripts/bench_cfg.py cfg1 | cut -c1-200
RGCN_OVERLAP_HUBS=$v python scripts/bench_cfg.py cfg3 | cut -c1-230
done
