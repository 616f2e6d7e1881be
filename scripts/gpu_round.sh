timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32.log 2>&1; echo "bench exit $?"; tail -c 1500 gpurun_out/bench_fp32.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --mode bf16 > gpurun_out/bench_bf16.log 2>&1; tail -c 600 gpurun_out/bench_bf16.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo ncu $?
