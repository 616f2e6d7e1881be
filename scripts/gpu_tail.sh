timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "aggregate or micro or rows or hub or full_size or golden" 2>&1 | tail -2
python scripts/bench_cfg.py cfg1 | cut -c90-200; python scripts/bench_cfg.py cfg3 | cut -c90-220
python bench.py --steps 20 --warmup 3 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 scripts/run_partitioned.py --nodes 1000000 --edges 40000000 --relations 30 --layers 3 --exchange nccl 2>&1 | tail -1 | cut -c1-400
