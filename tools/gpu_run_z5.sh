#!/bin/bash
# session Z5: windows with the relative-position product fused into the attention kernel
OUT=gpurun_out; mkdir -p $OUT
{
timeout 240 python tools/probes/win_attention_debug.py 2>&1 | head -8
echo "--- tests"
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "sam_attention" 2>&1 | tail -4
echo "--- probes"
for f in 0 1; do echo "RB200_ATTN_WIN_FUSE=$f"; for k in attn_sam_win attn_sam_win4; do RB200_ATTN_WIN_FUSE=$f timeout 120 python tools/kernel_probe.py $k 2>&1 | tail -1; done; done
echo "--- memcheck"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "sam_attention and bfloat16" > $OUT/z5_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY" $OUT/z5_memcheck.log | tail -3
echo "--- sam model tests"
timeout 600 python -m pytest tests -m gpu -x -q -k "sam or SAM or segment" 2>&1 | tail -3
echo "--- bench config 5"
timeout 600 python bench.py --config 5 > $OUT/z5_bench5.json 2> $OUT/z5_bench5.err; echo "rc=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/z5_bench5.json").read().strip().splitlines()[-1])
print(d["value"], d["unit"], d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d.get("gpu_launches"))
P
} > gpurun_out/z5_summary.txt 2>&1
cat gpurun_out/z5_summary.txt
