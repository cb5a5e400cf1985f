#!/bin/bash
# Round-2 GPU session P: SASS-level digests of the GEMM kernel (dominant N=1280 shape and the long GEGLU shape).
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
for probe in gemm gemm_geglu conv; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 1 -f -o $OUT/p_ncu_$probe python tools/kernel_probe.py $probe 3 > $OUT/p_ncu_$probe.log 2>&1
  python tools/ncu_source_digest.py $OUT/p_ncu_$probe.ncu-rep $OUT/p_digest_$probe.txt --top 45 > /dev/null 2>> $OUT/p_ncu_$probe.log
  python tools/ncu_summary.py $OUT/p_ncu_$probe.ncu-rep $OUT/p_summary_$probe.txt --what "$probe (session P)" >> $OUT/p_ncu_$probe.log 2>&1
  rm -f $OUT/p_ncu_$probe.ncu-rep
done
grep -A50 "^top" $OUT/p_digest_gemm_geglu.txt | head -60
