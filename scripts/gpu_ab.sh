for thr in 128 64 32; do for ord in degree none; do
echo "== thr $thr order $ord"; PRIMEKG_RGCN_HUB_THRESHOLD=$thr PRIMEKG_RGCN_ROW_ORDER=$ord python scripts/bench_agg.py 2>&1 | grep -v "fp32-out"
done; done
for thr in 128 64 32; do
echo "== step thr $thr"; PRIMEKG_RGCN_HUB_THRESHOLD=$thr python scripts/bench_cfg.py cfg2 | cut -c1-200;  PRIMEKG_RGCN_HUB_THRESHOLD=$thr python scripts/bench_cfg.py cfg3 | cut -c1-220
done
