#!/bin/bash
# Round-2 GPU session H: suite, probes, per-kernel ncu evidence SUMMARISED ON THE BOX (the .ncu-rep files are deleted:
# gpurun_out/ is capped at 64 MiB), bench configs 2-5, launch lists.
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT; P=$OUT/h_profiles; mkdir -p $P
timeout 1800 python -m pytest tests -q -m gpu -x > $OUT/h_t_all.log 2>&1; echo "gpu suite rc=$?" | tee -a $OUT/h_t_all.log
for probe in gemm gemm_res gemm_geglu gemm640_res conv conv320 gemm_kv attn attn4096 attn77 attn_sam_win attn_sam_global gn ln; do
  echo "--- $probe" >> $OUT/h_probes.txt
  timeout 120 python tools/kernel_probe.py $probe 20 >> $OUT/h_probes.txt 2>&1
done
cap() {  # name, kernel regex, probe, flops-or-bytes flag, value, shape, description
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o $OUT/h_ncu_$1 python tools/kernel_probe.py $3 3 > $OUT/h_ncu_$1.log 2>&1
  echo "rc=$?" >> $OUT/h_ncu_$1.log
  if [ -f $OUT/h_ncu_$1.ncu-rep ]; then
    EXTRA=""; [ -n "$6" ] && EXTRA="--shape $6 --json $P/r02_ncu_$1.json"
    python tools/ncu_summary.py $OUT/h_ncu_$1.ncu-rep $P/r02_ncu_$1.txt $4 $5 $EXTRA --what "$7" >> $OUT/h_ncu_$1.log 2>&1
    ncu -i $OUT/h_ncu_$1.ncu-rep --page source --csv 2>/dev/null | head -400 > $P/r02_ncu_$1.source_head.csv
    rm -f $OUT/h_ncu_$1.ncu-rep
  fi
}
cap gemm_dominant tc_gemm gemm --flops 53.687e9 16384,1280,1280 "dominant GEMM [16384,1280]x[1280,1280]^T bf16, pair mode, pipelined epilogue"
cap gemm_res tc_gemm gemm_res --flops 53.687e9 "" "same GEMM + bias + residual epilogue"
cap gemm640_res tc_gemm gemm640_res --flops 53.687e9 "" "[65536,640]x[640,640]^T + bias + residual"
cap gemm_geglu tc_gemm gemm_geglu --flops 429.5e9 "" "[16384,1280]x[10240,1280]^T + GEGLU epilogue"
cap conv_1280 tc_gemm conv --flops 483.2e9 "" "3x3 conv 1280->1280 @32x32 batch 16 (implicit GEMM)"
cap conv_320 tc_gemm conv320 --flops 483.2e9 "" "3x3 conv 320->320 @128x128 batch 16 (implicit GEMM)"
cap attention_v3_s1024 tc_sdpa2 attn --flops 85.9e9 "" "tc_sdpa2 (two query tiles, 16 softmax warps, two MMA issuers) B=16 H=20 S=1024 d=64"
cap attention_v3_s4096 tc_sdpa2 attn4096 --flops 687.2e9 "" "tc_sdpa2 B=16 H=10 S=4096 d=64"
cap attention_sk77 tc_sdpa_kernel attn77 --flops 6.46e9 "" "tc_sdpa_kernel text cross-attention Sq=1024 Sk=77 B=16 H=20 d=64"
cap sam_global tc_sdpa_kernel attn_sam_global --flops 85.9e9 "" "SAM global attention 4096 tokens, 16 heads, d=80 (+rel-pos bias)"
cap sam_relbias rel_bias attn_sam_win "" "" "" "SAM rel-pos bias tables (mma.sync) for 25 windows of 14x14"
cap gn_partial gn_partial gn --bytes 167.8e6 "" "GroupNorm statistics pass [16,320,128,128] bf16 (reads the tensor once)"
cap gn_apply gn_apply gn --bytes 335.5e6 "" "GroupNorm normalise+SiLU pass (read + write)"
cap ln layer_norm ln --bytes 83.9e6 "" "LayerNorm [16384,1280] bf16 (read + write)"
timeout 900 python bench.py --config 2 --steps 20 --warmup 5 > $OUT/h_bench2.json 2> $OUT/h_bench2.err; echo "rc=$?" >> $OUT/h_bench2.err
for cfg in 3 4 5; do
  timeout 900 python bench.py --config $cfg --steps 15 --warmup 3 --skip-cpu-baseline > $OUT/h_bench$cfg.json 2> $OUT/h_bench$cfg.err; echo "rc=$?" >> $OUT/h_bench$cfg.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/h_launches_cfg2.csv \
  python bench.py --profile-step --no-graph > $OUT/h_prof_step.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/h_launches_cfg5.csv \
  python bench.py --config 5 --profile-step --no-graph > $OUT/h_prof_step5.log 2>&1
tail -4 $OUT/h_t_all.log; grep -v "^$" $OUT/h_probes.txt | paste - - | cut -c1-150
for cfg in 2 3 4 5; do python - <<P
import json
try:
    d=json.loads(open("$OUT/h_bench$cfg.json").read().strip().splitlines()[-1])
    print($cfg, round(d["value"],3), d["unit"], round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"],3), "eager", d["gpu_eager_baseline"] and round(d["gpu_eager_baseline"].get("value",0),3), "launches/replay", d["config"]["launches_per_replay"])
except Exception as e: print($cfg, "failed", e)
P
done
du -sh $OUT
