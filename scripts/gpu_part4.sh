# N-GPU partitioned RGCN (destination-range shards), strong scaling on one fixed graph: fused peer-memory exchange vs NCCL
N=$1; NODES=$2; EDGES=$3
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 scripts/run_partitioned.py "${@:3}"; }
for ex in fused nccl; do
  run $N 29521 --nodes $NODES --edges $EDGES --relations 30 --layers 3 --exchange $ex --steps 5 > gpurun_out/part${N}_${ex}.log 2>&1; echo "N=$N $ex exit $?"; tail -1 gpurun_out/part${N}_${ex}.log | cut -c1-600
done
