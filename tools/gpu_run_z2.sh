#!/bin/bash
# session Z2: window attention kernel: error split, wider tests, per-kernel times, one full ncu capture
OUT=gpurun_out; mkdir -p $OUT
{
RB200_ATTN_WIN=1 timeout 180 python tools/probes/win_attention_debug.py 2>&1 | tail -8
echo "--- tests"
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "sam_attention" 2>&1 | tail -4
echo "--- launch list (batch 4)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/z2_launches.csv python tools/kernel_probe.py attn_sam_win4 3 > /dev/null 2>&1
python tools/summarize_launches.py $OUT/z2_launches.csv 2>&1 | head -8
echo "--- ncu full"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_sdpa_win -s 3 -c 1 -f -o $OUT/z2_ncu_win python tools/kernel_probe.py attn_sam_win4 3 > $OUT/z2_ncu.log 2>&1
python tools/ncu_summary.py $OUT/z2_ncu_win.ncu-rep $OUT/z2_ncu_win_summary.txt --flops 19.67e9 --bytes 246.6e6 --what "tc_sdpa_win SAM windows: 100 windows x 16 heads, 196 x 196, d=80 (batch 4)" >> $OUT/z2_ncu.log 2>&1
python tools/ncu_source_digest.py $OUT/z2_ncu_win.ncu-rep $OUT/z2_win_digest.txt --top 30 > /dev/null 2>> $OUT/z2_ncu.log
rm -f $OUT/z2_ncu_win.ncu-rep
cat $OUT/z2_ncu_win_summary.txt; sed -n 1,30p $OUT/z2_win_digest.txt
} > gpurun_out/z2_summary.txt 2>&1
cat gpurun_out/z2_summary.txt
