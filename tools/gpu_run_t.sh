#!/bin/bash
# Round-2 GPU session T: first-generation attention kernel after the dead LAZY variant was removed (SAM, head dim 128, dual
# beyond the short kernel's range), GroupNorm chunk-size experiment.
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_sam_and_adapters.py tests/test_full_size_gpu.py -q -x -m gpu -k "sdpa or sam or attention or SAM or cfg5" > $OUT/t_t.log 2>&1; echo "tests rc=$?" | tee -a $OUT/t_t.log
for pix in 128 256 512 64; do
  echo "=== RB200_GN_PIX=$pix" >> $OUT/t_gn.txt
  RB200_GN_PIX=$pix timeout 120 python tools/kernel_probe.py gn 20 >> $OUT/t_gn.txt 2>&1
done
for probe in attn_sam_win attn_sam_global; do timeout 120 python tools/kernel_probe.py $probe 20 >> $OUT/t_gn.txt 2>&1; done
tail -2 $OUT/t_t.log; cat $OUT/t_gn.txt
