for f in w z; do
PRIMEKG_RGCN_BASIS_FORM=$f ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_cfg3_$f.csv python scripts/bench_cfg.py cfg3 > gpurun_out/ncu_cfg3_$f.log 2>&1; echo "ncu $f $?"
done
