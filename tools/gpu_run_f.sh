#!/bin/bash
# Round-2 GPU session F: attention v3 (16 softmax warps) with a 2-stage K/V ring.
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "sdpa or cfg_euler" -x > $OUT/f_t_sdpa.log 2>&1
V2RC=$?; echo "attention tests rc=$V2RC" | tee -a $OUT/f_t_sdpa.log
for probe in attn attn4096 attn77 attn77_4096; do
  for cfg in "1 0" "1 1"; do
    set -- $cfg
    echo "--- $probe v2=$1 poly=$2" >> $OUT/f_probes.txt
    RB200_ATTN_V2=$1 RB200_ATTN_POLY=$2 timeout 120 python tools/kernel_probe.py $probe 20 >> $OUT/f_probes.txt 2>&1
  done
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_sdpa2 -s 3 -c 1 -f -o $OUT/f_attn3_1024 \
  python tools/kernel_probe.py attn 3 > $OUT/f_ncu_attn3.log 2>&1; echo "rc=$?" >> $OUT/f_ncu_attn3.log
[ $V2RC -eq 0 ] && timeout 600 python bench.py --config 2 --steps 15 --warmup 3 --skip-cpu-baseline --skip-eager-baseline > $OUT/f_bench2.json 2> $OUT/f_bench2.err
tail -3 $OUT/f_t_sdpa.log; grep -v "^$" $OUT/f_probes.txt | tail -20; python -c "
import json; d=json.loads(open('$OUT/f_bench2.json').read().strip().splitlines()[-1]); print('bench2', d['value'], d['ms_per_step'])"
