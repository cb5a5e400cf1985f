#!/bin/bash
# session Z4: SAM on the new attention kernel end to end: parity tests, memcheck, bench config 5 A/B
OUT=gpurun_out; mkdir -p $OUT
{
echo "--- sam tests"
timeout 600 python -m pytest tests -m gpu -x -q -k "sam or SAM or segment" 2>&1 | tail -4
echo "--- memcheck"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "sam_attention and bfloat16" > $OUT/z4_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY" $OUT/z4_memcheck.log | tail -3
echo "--- bench config 5"
RB200_ATTN_WIN=0 timeout 600 python bench.py --config 5 > $OUT/z4_bench5_old.json 2> $OUT/z4_bench5_old.err; echo "old rc=$?"
timeout 600 python bench.py --config 5 > $OUT/z4_bench5_new.json 2> $OUT/z4_bench5_new.err; echo "new rc=$?"
python - <<'P'
import json
for t in ("old","new"):
    try:
        d=json.loads(open(f"gpurun_out/z4_bench5_{t}.json").read().strip().splitlines()[-1])
        print(t, d["value"], d["unit"], d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"], "launches", d.get("gpu_launches"))
    except Exception as e: print(t, "unreadable", e)
P
} > gpurun_out/z4_summary.txt 2>&1
cat gpurun_out/z4_summary.txt
