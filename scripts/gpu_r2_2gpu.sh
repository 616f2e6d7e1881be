set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_fused_dist_2gpu.py -m gpu -q -x > gpurun_out/pytest_2gpu.log 2>&1; echo "2gpu tests exit $?"; tail -15 gpurun_out/pytest_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.log 2>gpurun_out/bench_2gpu.err; echo "bench 2gpu $?"; tail -c 400 gpurun_out/bench_2gpu.err
RGCN_DP_ALLREDUCE=nccl RGCN_PIPELINE=0 timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu_nccl_nopipe.log 2>&1; echo "bench 2gpu nccl/nopipe $?"
RGCN_PIPELINE=1 timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu_pipe1.log 2>&1; echo "bench 2gpu pipe1 $?"
python - <<'PY'
import json
for f in ("bench_2gpu","bench_2gpu_nccl_nopipe","bench_2gpu_pipe1"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.log"%f).read().strip().splitlines() if l.startswith("{")][-1]);print(f,d["ms_per_step"],d["value"],d["e2e"]["value"],d["config"]["parallelism"][:90]); p=d.get("partitioned") or {}; print("  part", p.get("ms_per_step"), p.get("nccl_exchange_ms_per_step"), p.get("equals_single_gpu"), p.get("error"), p.get("fused_error"))
    except Exception as e: print(f, "ERR", e)
PY
