#!/bin/bash
# Round-2 GPU session O: short-key attention with the deferred epilogue: parity + probes + ncu summary.
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "sdpa" > $OUT/o_t_sdpa.log 2>&1; echo "sdpa tests rc=$?" | tee -a $OUT/o_t_sdpa.log
for probe in attn77 attn77_4096; do timeout 120 python tools/kernel_probe.py $probe 20 >> $OUT/o_probes.txt 2>&1; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_sdpa_short -s 3 -c 1 -f -o $OUT/o_ncu_short python tools/kernel_probe.py attn77 3 > $OUT/o_ncu_short.log 2>&1
python tools/ncu_summary.py $OUT/o_ncu_short.ncu-rep $OUT/o_ncu_short_summary.txt --flops 6.46e9 --bytes 88.1e6 --what "tc_sdpa_short text cross-attention B=16 H=20 Sq=1024 Sk=77 d=64" >> $OUT/o_ncu_short.log 2>&1
python tools/ncu_source_digest.py $OUT/o_ncu_short.ncu-rep $OUT/o_short_digest.txt --top 30 > /dev/null 2>> $OUT/o_ncu_short.log
rm -f $OUT/o_ncu_short.ncu-rep
tail -2 $OUT/o_t_sdpa.log; cat $OUT/o_probes.txt; head -22 $OUT/o_ncu_short_summary.txt; sed -n 1,24p $OUT/o_short_digest.txt; grep -A14 "^top" $OUT/o_short_digest.txt
