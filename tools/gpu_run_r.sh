#!/bin/bash
# Round-2 GPU session R: IP-Adapter (dual key / value set) form of the short-key attention kernel.
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "sdpa" > $OUT/r_t_sdpa.log 2>&1; echo "sdpa tests rc=$?" | tee -a $OUT/r_t_sdpa.log
for v in 1 0; do
  echo "=== RB200_ATTN_SHORT_DUAL=$v" >> $OUT/r_probes.txt
  RB200_ATTN_SHORT_DUAL=$v timeout 120 python tools/kernel_probe.py attn77_dual 20 >> $OUT/r_probes.txt 2>&1
done
timeout 900 python -m pytest tests/test_sam_and_adapters.py tests/test_full_size_gpu.py -q -m gpu -x > $OUT/r_t_models.log 2>&1; echo "model tests rc=$?" | tee -a $OUT/r_t_models.log
timeout 600 python bench.py --config 3 --steps 15 --warmup 3 --skip-cpu-baseline > $OUT/r_bench3.json 2> $OUT/r_bench3.err
timeout 600 python bench.py --config 2 --steps 15 --warmup 3 --skip-cpu-baseline --skip-eager-baseline > $OUT/r_bench2.json 2> $OUT/r_bench2.err
tail -2 $OUT/r_t_sdpa.log; cat $OUT/r_probes.txt; tail -2 $OUT/r_t_models.log
python - <<P
import json
for c in (3, 2):
    try:
        d=json.loads(open("$OUT/r_bench%d.json" % c).read().strip().splitlines()[-1]); print(c, round(d["value"],3), d["ms_per_step"])
    except Exception as e: print(c, "failed", e)
P
