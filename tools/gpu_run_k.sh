#!/bin/bash
# Round-2 GPU session K: attention after the instruction diet (packed fp32 pairs, hoisted shared addresses) with the turnstile
# on by default; parity of every attention test; SASS digest; new feature tests (text encoders, avg pool).
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
for v in 2 0; do
  echo "=== STAGGER=$v" >> $OUT/k_attn.txt
  RB200_ATTN_STAGGER=$v timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "sdpa" 2>&1 | tail -1 >> $OUT/k_attn.txt
  for probe in attn attn4096; do
    RB200_ATTN_STAGGER=$v timeout 120 python tools/kernel_probe.py $probe 20 >> $OUT/k_attn.txt 2>&1
  done
done
timeout 900 python -m pytest tests/test_encoders.py tests/test_full_size_gpu.py -q -m gpu -x > $OUT/k_t.log 2>&1; echo "tests rc=$?" | tee -a $OUT/k_t.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_sdpa2 -s 3 -c 1 -f -o $OUT/k_ncu_attn python tools/kernel_probe.py attn4096 3 > $OUT/k_ncu_attn.log 2>&1
python tools/ncu_source_digest.py $OUT/k_ncu_attn.ncu-rep $OUT/k_attn_source_digest.txt --top 60 > /dev/null 2>> $OUT/k_ncu_attn.log
python tools/ncu_summary.py $OUT/k_ncu_attn.ncu-rep $OUT/k_ncu_attn_summary.txt --flops 687.2e9 --what "tc_sdpa2 S=4096: turnstile + packed fp32 pairs + hoisted addresses (session K)" >> $OUT/k_ncu_attn.log 2>&1
rm -f $OUT/k_ncu_attn.ncu-rep
cat $OUT/k_attn.txt; tail -3 $OUT/k_t.log; head -48 $OUT/k_attn_source_digest.txt; grep -E "duration|tensor pipe|XU|issue|warp-state" $OUT/k_ncu_attn_summary.txt
