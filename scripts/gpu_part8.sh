# cfg5 of BASELINE.json: 10 M nodes / 400 M edges / 30 relations / 3 layers on 8 B200, fused peer-memory exchange, then NCCL
run() { timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 scripts/run_partitioned.py "${@:2}"; }
run 29531 --nodes 10000000 --edges 400000000 --relations 30 --layers 3 --exchange fused --steps 3 > gpurun_out/part8_cfg5_fused.log 2>&1; echo "cfg5 fused exit $?"; tail -1 gpurun_out/part8_cfg5_fused.log | cut -c1-700
nvidia-smi --query-gpu=memory.used --format=csv,noheader | head -2
run 29532 --nodes 10000000 --edges 400000000 --relations 30 --layers 3 --exchange nccl --steps 3 > gpurun_out/part8_cfg5_nccl.log 2>&1; echo "cfg5 nccl exit $?"; tail -1 gpurun_out/part8_cfg5_nccl.log | cut -c1-700
