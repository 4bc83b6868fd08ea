for c in 74 148 222 296 444 592; do SSR_WGRAD_BATCH_CTAS=$c timeout 100 python scripts/bench_rdb_bwd.py 2>&1 | tail -1; done
SSR_CHAIN_TIMELINE=1 timeout 100 python scripts/chain_timeline.py 2>&1 | grep -E "^==|mean over|->|layer total"
