timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_gpu.log
for f in w z; do PRIMEKG_RGCN_BASIS_FORM=$f python scripts/bench_cfg.py cfg3 2>&1 | tail -2; done
