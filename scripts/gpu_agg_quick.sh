timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "aggregate or hub or basis or micro or golden" 2>&1 | tail -4
python scripts/bench_agg.py 2>&1 | grep -v "fp32-out"
python scripts/bench_cfg.py cfg2 | cut -c1-200; python scripts/bench_cfg.py cfg3 | cut -c1-220
