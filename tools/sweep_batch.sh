#!/bin/bash
# row-path batch size sweep (L2-resident batches vs large batches)
for mb in 0 131072 262144 1048576; do
  echo "max_batch=$mb"
  CPPROB_SIS_MAX_BATCH=$mb python tools/bench_configs.py 2>&1 | grep -E '"C3"|"C4"|C5 stats' | python tools/_fmt_cfg.py
done
