#!/bin/bash
# Same-box A/B of the GEMM cluster modes (RB200_GEMM_CLUSTER=2 multicast pair, =3 cta_group::2 pair):
# SDXL resident step time (stderr log line) and SAM images/s, alternating so that thermal drift shows.
for m in 2 3 2 3; do
  echo "MODE $m"
  RB200_GEMM_CLUSTER=$m timeout 300 python bench.py --steps 20 --skip-cpu-baseline --resident-only 2>&1 | grep "resident loop done" | tail -1
  RB200_GEMM_CLUSTER=$m timeout 100 python tools/bench_sam.py --batches 4 | cut -c1-110
done
