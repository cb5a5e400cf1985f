#!/bin/bash
# Round-2 GPU session V: two-threads-per-row variant of the short-key attention kernel (parity, probes split vs not), T2I graph test.
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "sdpa" > $OUT/v_t_sdpa.log 2>&1; echo "sdpa tests rc=$?" | tee -a $OUT/v_t_sdpa.log
for v in 1 0; do
  echo "=== RB200_ATTN_SHORT_SPLIT=$v" >> $OUT/v_probes.txt
  for probe in attn77 attn77_4096 attn77_dual; do
    RB200_ATTN_SHORT_SPLIT=$v timeout 120 python tools/kernel_probe.py $probe 20 >> $OUT/v_probes.txt 2>&1
  done
done
timeout 900 python -m pytest tests/test_t2i_adapter.py tests/test_sam_and_adapters.py tests/test_full_size_gpu.py -q -m gpu -x > $OUT/v_t_models.log 2>&1; echo "model tests rc=$?" | tee -a $OUT/v_t_models.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 77 --print-limit 20 python -m pytest tests/test_kernels_gpu.py -k "short_keys and not 1024 and not 4096" -q -x -m gpu > $OUT/v_mem.log 2>&1; echo "memcheck rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' $OUT/v_mem.log | tr '\n' ' ')" | tee -a $OUT/v_t_models.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_sdpa_short -s 3 -c 1 -f -o $OUT/v_ncu_short python tools/kernel_probe.py attn77 3 > $OUT/v_ncu_short.log 2>&1
python tools/ncu_summary.py $OUT/v_ncu_short.ncu-rep $OUT/v_ncu_short_summary.txt --flops 6.46e9 --bytes 88.1e6 --what "tc_sdpa_short (two threads per row) text cross-attention B=16 H=20 Sq=1024 Sk=77 d=64" >> $OUT/v_ncu_short.log 2>&1
python tools/ncu_source_digest.py $OUT/v_ncu_short.ncu-rep $OUT/v_short_digest.txt --top 30 > /dev/null 2>> $OUT/v_ncu_short.log
rm -f $OUT/v_ncu_short.ncu-rep
tail -2 $OUT/v_t_sdpa.log; cat $OUT/v_probes.txt; tail -3 $OUT/v_t_models.log; head -20 $OUT/v_ncu_short_summary.txt
