cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_8gpu_r2.log 2> gpurun_out/bench_8gpu_r2.err; echo "rc $?"
tail -c 600 gpurun_out/bench_8gpu_r2.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_8gpu_r2.log'):
    if l.startswith('{'):
        d=json.loads(l); print('N', d['n_gpus'], 'ms', d['ms_per_step'], 'value', d['value']); p=d.get('partitioned'); print(json.dumps(p)[:900] if p else None)
PY
