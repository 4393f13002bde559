#!/bin/bash
# Sweep of the persistent path's shape for the latency-bound sizes: cluster size x gather mode
# (x CTAs of the streaming kernel when the resident kernel is switched off).
cd "$(dirname "$0")/.."
for cl in 1 2 4 8; do
  for dm in 0 100000; do
    NLO_CLUSTER=$cl NLO_DIRECT_MAX=$dm timeout 120 python scripts/latency_small.py "$@" 2>&1 | tail -1
  done
done
