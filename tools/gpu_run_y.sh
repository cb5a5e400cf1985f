#!/bin/bash
# Round-2 GPU session Y: four-CTA multicast clusters in the GEMM (RB200_GEMM_CLUSTER=4) vs the shipped CTA-pair mode (3).
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
RB200_GEMM_CLUSTER=4 timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "linear or conv or geglu" > $OUT/y_t.log 2>&1; echo "mode 4 tests rc=$? $(tail -1 $OUT/y_t.log)" | tee -a $OUT/y_probes.txt
for m in 4 3 2; do
  echo "=== RB200_GEMM_CLUSTER=$m" >> $OUT/y_probes.txt
  for probe in gemm_geglu gemm gemm_res conv conv320 gemm640_res; do
    RB200_GEMM_CLUSTER=$m timeout 120 python tools/kernel_probe.py $probe 20 >> $OUT/y_probes.txt 2>&1
  done
done
cat $OUT/y_probes.txt
