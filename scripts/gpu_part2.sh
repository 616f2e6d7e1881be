# 2-GPU: fused (peer memory) vs nccl exchange, correctness then timing
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 scripts/run_partitioned.py "${@:3}"; }
for ex in fused nccl; do
  run 2 29511 --nodes 200000 --edges 4000000 --relations 30 --layers 3 --check --exchange $ex > gpurun_out/part2_check_$ex.log 2>&1; echo "check $ex exit $?"; tail -2 gpurun_out/part2_check_$ex.log | cut -c1-900
done
for ex in fused nccl; do
  run 2 29512 --nodes 1000000 --edges 40000000 --relations 30 --layers 3 --exchange $ex > gpurun_out/part2_time_$ex.log 2>&1; echo "time $ex exit $?"; tail -1 gpurun_out/part2_time_$ex.log | cut -c1-700
done
PRIMEKG_RGCN_PEER=ipc run 2 29513 --nodes 1000000 --edges 40000000 --relations 30 --layers 3 --exchange fused > gpurun_out/part2_time_fused_ipc.log 2>&1; echo "time fused ipc exit $?"; tail -1 gpurun_out/part2_time_fused_ipc.log | cut -c1-700
