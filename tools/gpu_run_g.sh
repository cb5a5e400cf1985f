#!/bin/bash
# Round-2 GPU session G: full suite, bench configs 2-5 (complete lines), per-kernel ncu evidence, launch lists.
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "sdpa" -x > $OUT/g_t_sdpa.log 2>&1
V2RC=$?; echo "attention tests rc=$V2RC" | tee -a $OUT/g_t_sdpa.log
[ $V2RC -ne 0 ] && export RB200_ATTN_V2=0
timeout 1800 python -m pytest tests -q -m gpu -s > $OUT/g_t_all.log 2>&1; echo "gpu suite rc=$?" | tee -a $OUT/g_t_all.log
for probe in attn attn4096 attn77 attn77_4096 attn_sam_win attn_sam_global gemm gemm_res gemm_geglu gemm640_res conv conv320 gemm_kv gn ln; do
  echo "--- $probe" >> $OUT/g_probes.txt
  timeout 120 python tools/kernel_probe.py $probe 20 >> $OUT/g_probes.txt 2>&1
done
cap() {  # name, kernel regex, probe
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o $OUT/g_ncu_$1 python tools/kernel_probe.py $3 3 > $OUT/g_ncu_$1.log 2>&1
  echo "rc=$?" >> $OUT/g_ncu_$1.log
}
cap attn_v3_1024 tc_sdpa2 attn
cap attn_v3_4096 tc_sdpa2 attn4096
cap attn_sk77 tc_sdpa_kernel attn77
cap sam_global tc_sdpa_kernel attn_sam_global
cap sam_relbias rel_bias attn_sam_win
cap conv_1280 tc_gemm conv
cap gemm_geglu tc_gemm gemm_geglu
cap gn_partial gn_partial gn
cap gn_apply gn_apply gn
cap ln layer_norm ln
timeout 900 python bench.py --config 2 --steps 20 --warmup 5 > $OUT/g_bench2.json 2> $OUT/g_bench2.err; echo "rc=$?" >> $OUT/g_bench2.err
for cfg in 3 4 5; do
  timeout 900 python bench.py --config $cfg --steps 15 --warmup 3 --skip-cpu-baseline > $OUT/g_bench$cfg.json 2> $OUT/g_bench$cfg.err; echo "rc=$?" >> $OUT/g_bench$cfg.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/g_launches_cfg2.csv \
  python bench.py --profile-step --no-graph > $OUT/g_prof_step.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/g_launches_cfg5.csv \
  python bench.py --config 5 --profile-step --no-graph > $OUT/g_prof_step5.log 2>&1
tail -4 $OUT/g_t_all.log; grep -v "^$" $OUT/g_probes.txt
for cfg in 2 3 4 5; do python - <<P
import json
try:
    d=json.loads(open("$OUT/g_bench$cfg.json").read().strip().splitlines()[-1])
    print($cfg, round(d["value"],3), d["unit"], round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"],3), "eager", d["gpu_eager_baseline"] and round(d["gpu_eager_baseline"].get("value",0),3), "launches/replay", d["config"]["launches_per_replay"])
except Exception as e: print($cfg, "failed", e)
P
done
