#!/bin/bash
# Round-2 GPU session U: compute-sanitizer memcheck over the kernels added this round (small shapes), then the T2I graph test.
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
run() {  # name, pytest args...
  name=$1; shift
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 --print-limit 20 python -m pytest "$@" -q -x -m gpu > $OUT/u_mem_$name.log 2>&1
  echo "$name rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' $OUT/u_mem_$name.log | tr '\n' ' ')" | tee -a $OUT/u_summary.txt
}
run style tests/test_style_aligned.py -k "kernel"
run pool tests/test_t2i_adapter.py -k "avg_pool"
run gnfixed tests/test_models_golden.py -k "group_norm_fixed"
run probs tests/test_sag.py -k "attention_probs and not 1024"
run glue tests/test_kernels_gpu.py -k "cfg_euler"
run short tests/test_kernels_gpu.py -k "short_keys and not 1024 and not 4096"
run attn2 tests/test_kernels_gpu.py -k "test_sdpa and bfloat16 and not production and not short and not growing"
timeout 600 python -m pytest tests/test_t2i_adapter.py tests/test_fluxion_api.py -q -x -m gpu > $OUT/u_t2i.log 2>&1; echo "t2i graph rc=$? $(tail -1 $OUT/u_t2i.log)" | tee -a $OUT/u_summary.txt
cat $OUT/u_summary.txt
