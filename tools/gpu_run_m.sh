#!/bin/bash
# Round-2 GPU session M: the single-pass short-key attention kernel (parity, probes old vs new), StyleAligned GPU tests.
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "sdpa" > $OUT/m_t_sdpa.log 2>&1; echo "sdpa tests rc=$?" | tee -a $OUT/m_t_sdpa.log
for v in 1 0; do
  echo "=== RB200_ATTN_SHORT=$v" >> $OUT/m_probes.txt
  for probe in attn77 attn77_4096; do
    RB200_ATTN_SHORT=$v timeout 120 python tools/kernel_probe.py $probe 20 >> $OUT/m_probes.txt 2>&1
  done
done
timeout 900 python -m pytest tests/test_style_aligned.py tests/test_sam_and_adapters.py tests/test_models_golden.py tests/test_full_size_gpu.py -q -m gpu -x > $OUT/m_t_models.log 2>&1; echo "model tests rc=$?" | tee -a $OUT/m_t_models.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_sdpa_short -s 3 -c 1 -f -o $OUT/m_ncu_short python tools/kernel_probe.py attn77 3 > $OUT/m_ncu_short.log 2>&1
python tools/ncu_summary.py $OUT/m_ncu_short.ncu-rep $OUT/m_ncu_short_summary.txt --flops 6.46e9 --bytes 88.1e6 --what "tc_sdpa_short text cross-attention B=16 H=20 Sq=1024 Sk=77 d=64" >> $OUT/m_ncu_short.log 2>&1
python tools/ncu_source_digest.py $OUT/m_ncu_short.ncu-rep $OUT/m_short_digest.txt --top 40 > /dev/null 2>> $OUT/m_ncu_short.log
rm -f $OUT/m_ncu_short.ncu-rep
timeout 600 python bench.py --config 2 --steps 20 --warmup 5 --skip-cpu-baseline > $OUT/m_bench2.json 2> $OUT/m_bench2.err
tail -3 $OUT/m_t_sdpa.log; cat $OUT/m_probes.txt; tail -3 $OUT/m_t_models.log; cat $OUT/m_ncu_short_summary.txt | head -24
python - <<P
import json
d=json.loads(open("$OUT/m_bench2.json").read().strip().splitlines()[-1]); print(round(d["value"],3), d["ms_per_step"])
P
