#!/bin/bash
# Round-2 GPU session S: the whole GPU suite on the current tree, attention ncu summaries of the shipped kernel, bench
# configs 2-5, launch list of one step.
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT; P=$OUT/s_profiles; mkdir -p $P
timeout 1800 python -m pytest tests -q -m gpu -x > $OUT/s_t_all.log 2>&1; echo "gpu suite rc=$?" | tee -a $OUT/s_t_all.log
for probe in attn attn4096 attn77 attn77_4096 attn77_dual attn_sam_global attn_sam_win gemm gemm_res conv gn ln; do
  timeout 120 python tools/kernel_probe.py $probe 20 >> $OUT/s_probes.txt 2>&1
done
cap() {
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o $OUT/s_ncu_$1 python tools/kernel_probe.py $3 3 > $OUT/s_ncu_$1.log 2>&1
  if [ -f $OUT/s_ncu_$1.ncu-rep ]; then
    python tools/ncu_summary.py $OUT/s_ncu_$1.ncu-rep $P/r02_ncu_$1.txt --flops $4 --what "$5" >> $OUT/s_ncu_$1.log 2>&1
    python tools/ncu_source_digest.py $OUT/s_ncu_$1.ncu-rep $P/r02_ncu_$1.sass_digest.txt --top 40 > /dev/null 2>> $OUT/s_ncu_$1.log
    rm -f $OUT/s_ncu_$1.ncu-rep
  fi
}
cap attention_v3_s1024 tc_sdpa2 attn 85.9e9 "tc_sdpa2 (shipped: two query tiles, 16 softmax warps, two MMA issuers, turnstile, packed fp32 pairs) B=16 H=20 S=1024 d=64"
cap attention_v3_s4096 tc_sdpa2 attn4096 687.2e9 "tc_sdpa2 (shipped) B=16 H=10 S=4096 d=64"
cap attention_short_dual tc_sdpa_short attn77_dual 6.8e9 "tc_sdpa_short, IP-Adapter form: 77 text + 4 image tokens, B=16 H=20 Sq=1024 d=64"
timeout 900 python bench.py --config 2 --steps 20 --warmup 5 > $OUT/s_bench2.json 2> $OUT/s_bench2.err; echo "rc=$?" >> $OUT/s_bench2.err
for cfg in 3 4 5; do
  timeout 900 python bench.py --config $cfg --steps 15 --warmup 3 --skip-cpu-baseline > $OUT/s_bench$cfg.json 2> $OUT/s_bench$cfg.err; echo "rc=$?" >> $OUT/s_bench$cfg.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/s_launches_cfg2.csv \
  python bench.py --profile-step --no-graph > $OUT/s_prof_step.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/s_launches_cfg3.csv \
  python bench.py --config 3 --profile-step --no-graph > $OUT/s_prof_step3.log 2>&1
tail -4 $OUT/s_t_all.log; cat $OUT/s_probes.txt
for cfg in 2 3 4 5; do python - <<P
import json
try:
    d=json.loads(open("$OUT/s_bench$cfg.json").read().strip().splitlines()[-1])
    print($cfg, round(d["value"],3), d["unit"], round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"],3), "eager", d["gpu_eager_baseline"] and round(d["gpu_eager_baseline"].get("value",0),3), "roofline", round(d["roofline"]["frac"],3), d["roofline"].get("after_step_loops",{}).get("frac"), "launches/replay", d["config"]["launches_per_replay"])
except Exception as e: print($cfg, "failed", e)
P
done
du -sh $OUT
