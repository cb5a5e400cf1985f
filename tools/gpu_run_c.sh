#!/bin/bash
# Round-2 GPU session C: attention v2 with the MUFU turnstile between the two softmax groups.
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "sdpa" -x > $OUT/c_t_sdpa_v2.log 2>&1
V2RC=$?; echo "attention v2 tests rc=$V2RC" | tee -a $OUT/c_t_sdpa_v2.log
for probe in attn attn4096 attn77 attn77_4096; do
  for cfg in "1 0" "1 1"; do
    set -- $cfg
    echo "--- $probe v2=$1 poly=$2" >> $OUT/c_probes.txt
    RB200_ATTN_V2=$1 RB200_ATTN_POLY=$2 timeout 120 python tools/kernel_probe.py $probe 20 >> $OUT/c_probes.txt 2>&1
  done
done
RB200_ATTN_V2=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_sdpa2 -s 3 -c 1 -f -o $OUT/c_attn2_1024 \
  python tools/kernel_probe.py attn 3 > $OUT/c_ncu_attn2.log 2>&1; echo "rc=$?" >> $OUT/c_ncu_attn2.log
timeout 600 python bench.py --config 2 --steps 15 --warmup 3 --skip-cpu-baseline --skip-eager-baseline > $OUT/c_bench2.json 2> $OUT/c_bench2.err
tail -3 $OUT/c_t_sdpa_v2.log; grep -v "^$" $OUT/c_probes.txt; python -c "
import json; d=json.loads(open('$OUT/c_bench2.json').read().strip().splitlines()[-1]); print('bench2', d['value'], d['ms_per_step'])"
