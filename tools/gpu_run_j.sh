#!/bin/bash
# Round-2 GPU session J: attention experiments (exponential offload, de-phasing the two query tiles), SASS-level digest of the
# attention kernel, bench config 2 with the dominant kernel timed before / after the loops.
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
for v in "0 0" "1 0" "2 0" "0 1" "0 2" "1 1" "2 2" "1 2"; do
  set -- $v
  echo "=== POLY=$1 STAGGER=$2" >> $OUT/j_attn.txt
  RB200_ATTN_POLY=$1 RB200_ATTN_STAGGER=$2 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "sdpa" 2>&1 | tail -1 >> $OUT/j_attn.txt
  for probe in attn attn4096; do
    RB200_ATTN_POLY=$1 RB200_ATTN_STAGGER=$2 timeout 120 python tools/kernel_probe.py $probe 20 >> $OUT/j_attn.txt 2>&1
  done
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_sdpa2 -s 3 -c 1 -f -o $OUT/j_ncu_attn python tools/kernel_probe.py attn4096 3 > $OUT/j_ncu_attn.log 2>&1
python tools/ncu_source_digest.py $OUT/j_ncu_attn.ncu-rep $OUT/j_attn_source_digest.txt --top 90 > /dev/null 2>> $OUT/j_ncu_attn.log
python tools/ncu_summary.py $OUT/j_ncu_attn.ncu-rep $OUT/j_ncu_attn_summary.txt --flops 687.2e9 --what "tc_sdpa2 S=4096 (session J)" >> $OUT/j_ncu_attn.log 2>&1
rm -f $OUT/j_ncu_attn.ncu-rep
timeout 900 python bench.py --config 2 --steps 20 --warmup 5 > $OUT/j_bench2.json 2> $OUT/j_bench2.err; echo "rc=$?" >> $OUT/j_bench2.err
cat $OUT/j_attn.txt; head -60 $OUT/j_attn_source_digest.txt
python - <<P
import json
d=json.loads(open("$OUT/j_bench2.json").read().strip().splitlines()[-1])
print(round(d["value"],3), d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["ms_per_launch"], d["roofline"].get("after_step_loops"))
P
