#!/bin/bash
# Round-2 GPU session B: fixed TMEM probe, attention v2 with early S issue (tests, probes, ncu --set full with source),
# ncu of the dominant GEMMs, full GPU suite (merged LoRA path), bench configs 3 / 4 / 5.
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
echo "== tmem probe"; timeout 120 tools/probes/tmem_probe > $OUT/b_tmem_probe.txt 2>&1; echo "rc=$?" >> $OUT/b_tmem_probe.txt
echo "== attention v2 unit tests"
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "sdpa" -x > $OUT/b_t_sdpa_v2.log 2>&1
V2RC=$?; echo "attention v2 tests rc=$V2RC" | tee -a $OUT/b_t_sdpa_v2.log
[ $V2RC -ne 0 ] && export RB200_ATTN_V2=0
echo "== attention probes"
for probe in attn attn4096 attn77 attn77_4096; do
  for cfg in "0 0" "1 0" "1 1"; do
    set -- $cfg
    echo "--- $probe v2=$1 poly=$2" >> $OUT/b_probes.txt
    RB200_ATTN_V2=$1 RB200_ATTN_POLY=$2 timeout 120 python tools/kernel_probe.py $probe 20 >> $OUT/b_probes.txt 2>&1
  done
done
echo "== ncu attention v2 (S = 1024)"
RB200_ATTN_V2=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_sdpa2 -s 3 -c 1 -f -o $OUT/b_attn2_1024 \
  python tools/kernel_probe.py attn 3 > $OUT/b_ncu_attn2.log 2>&1; echo "rc=$?" >> $OUT/b_ncu_attn2.log
echo "== ncu attention v2 (Sk = 77)"
RB200_ATTN_V2=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_sdpa2 -s 3 -c 1 -f -o $OUT/b_attn2_77 \
  python tools/kernel_probe.py attn77 3 > $OUT/b_ncu_attn2_77.log 2>&1; echo "rc=$?" >> $OUT/b_ncu_attn2_77.log
echo "== ncu gemm 1280 (+ residual) and gemm 640 (+ residual)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 1 -f -o $OUT/b_gemm_res \
  python tools/kernel_probe.py gemm_res 3 > $OUT/b_ncu_gemm_res.log 2>&1; echo "rc=$?" >> $OUT/b_ncu_gemm_res.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 1 -f -o $OUT/b_gemm640_res \
  python tools/kernel_probe.py gemm640_res 3 > $OUT/b_ncu_gemm640_res.log 2>&1; echo "rc=$?" >> $OUT/b_ncu_gemm640_res.log
echo "== full gpu suite"
timeout 1500 python -m pytest tests -q -m gpu -s > $OUT/b_t_all.log 2>&1; echo "gpu suite rc=$?" | tee -a $OUT/b_t_all.log
for cfg in 2 3 4 5; do
  echo "== bench config $cfg"
  timeout 900 python bench.py --config $cfg --steps 15 --warmup 3 --skip-cpu-baseline > $OUT/b_bench$cfg.json 2> $OUT/b_bench$cfg.err; echo "bench rc=$?" >> $OUT/b_bench$cfg.err
done
RB200_LORA_MERGE=0 timeout 900 python bench.py --config 3 --steps 15 --warmup 3 --skip-cpu-baseline --skip-eager-baseline > $OUT/b_bench3_nomerge.json 2> $OUT/b_bench3_nomerge.err
tail -4 $OUT/b_t_all.log; cat $OUT/b_tmem_probe.txt; grep -v "^$" $OUT/b_probes.txt; for cfg in 2 3 4 5; do python - <<P
import json
try:
    d=json.loads(open("$OUT/b_bench$cfg.json").read().strip().splitlines()[-1])
    print($cfg, round(d["value"],3), d["unit"], round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"],3), "eager", d["gpu_eager_baseline"], "launches/replay", d["config"]["launches_per_replay"])
except Exception as e: print($cfg, "failed", e)
P
done; tail -c 600 $OUT/b_bench3_nomerge.json
