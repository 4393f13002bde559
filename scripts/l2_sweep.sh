#!/bin/bash
# L2 residency sweep: microseconds per iteration vs MB of the scan pinned in L2 (evict_last).
for keep in 0 32 48 64 80 96 110; do
  for n in 600000 1000000 1500000 2000000 4000000; do
    echo -n "keep=${keep}MB "; NLO_L2_KEEP_MB=$keep python scripts/one_solve.py $n ndt3 40
  done
done
