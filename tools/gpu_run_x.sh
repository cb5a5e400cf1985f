#!/bin/bash
# Round-2 GPU session X: SAM ViT-H encoder batch sweep (BASELINE config 5: batch 1-64), Perceiver GPU test.
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_encoders.py -q -x -m gpu -k "perceiver" > $OUT/x_t.log 2>&1; echo "perceiver rc=$? $(tail -1 $OUT/x_t.log)" | tee -a $OUT/x_sweep.txt
for b in 1 2 4 8 16 32 64; do
  timeout 600 python bench.py --config 5 --latent-batch $b --steps 6 --warmup 3 --skip-cpu-baseline --skip-eager-baseline > $OUT/x_sam_b$b.json 2> $OUT/x_sam_b$b.err
  python - <<P | tee -a $OUT/x_sweep.txt
import json
try:
    d=json.loads(open("$OUT/x_sam_b$b.json").read().strip().splitlines()[-1]); print("batch", $b, round(d["value"],2), d["unit"], round(d["ms_per_step"],2), "ms per batch; e2e", round(d["e2e"]["value"],2))
except Exception as e: print("batch", $b, "failed", e)
P
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_sdpa_kernel -s 3 -c 1 -f -o $OUT/x_ncu_samwin python tools/kernel_probe.py attn_sam_win 3 > $OUT/x_ncu_samwin.log 2>&1
python tools/ncu_summary.py $OUT/x_ncu_samwin.ncu-rep $OUT/x_ncu_samwin_summary.txt --flops 4.9e9 --what "tc_sdpa_kernel<HD=128, BIAS> SAM window attention: 25 windows of 14x14 tokens, 16 heads, d=80" >> $OUT/x_ncu_samwin.log 2>&1
python tools/ncu_source_digest.py $OUT/x_ncu_samwin.ncu-rep $OUT/x_samwin_digest.txt --top 40 > /dev/null 2>> $OUT/x_ncu_samwin.log
rm -f $OUT/x_ncu_samwin.ncu-rep
head -22 $OUT/x_ncu_samwin_summary.txt; grep -A24 "^top" $OUT/x_samwin_digest.txt
