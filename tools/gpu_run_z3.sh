#!/bin/bash
# session Z3: the global (64-wide map) geometry on the window kernel
OUT=gpurun_out; mkdir -p $OUT
{
timeout 240 python tools/probes/win_attention_debug.py 2>&1 | tail -11
echo "--- tests"
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "sam_attention" 2>&1 | tail -4
echo "--- probes"
for w in 1 2; do for k in attn_sam_global attn_sam_global4; do echo "RB200_ATTN_WIN=$w"; RB200_ATTN_WIN=$w timeout 120 python tools/kernel_probe.py $k 2>&1 | tail -1; done; done
echo "--- launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/z3_launches.csv python tools/kernel_probe.py attn_sam_global4 3 > /dev/null 2>&1
python tools/summarize_launches.py $OUT/z3_launches.csv 2>&1 | head -8
} > gpurun_out/z3_summary.txt 2>&1
cat gpurun_out/z3_summary.txt
