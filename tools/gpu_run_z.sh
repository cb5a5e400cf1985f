#!/bin/bash
# session Z: first run of the SAM window attention kernel (tc_attention_win.cu)
mkdir -p gpurun_out
{
RB200_ATTN_WIN=1 timeout 180 python tools/probes/win_attention_debug.py 2>&1 | tail -12
echo "--- old kernel"
RB200_ATTN_WIN=0 timeout 180 python tools/probes/win_attention_debug.py 2>&1 | tail -8
echo "--- tests"
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "sam_attention" 2>&1 | tail -4
echo "--- probes"
for w in 0 1; do for k in attn_sam_win attn_sam_win4; do RB200_ATTN_WIN=$w timeout 120 python tools/kernel_probe.py $k 2>&1 | tail -1; done; done
} > gpurun_out/z_summary.txt 2>&1
cat gpurun_out/z_summary.txt
