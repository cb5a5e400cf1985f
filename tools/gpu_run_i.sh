#!/bin/bash
# Round-2 GPU session I: new feature tests (SAG, tiled VAE), GEMM / conv ncu evidence re-captured on the shipped epilogue,
# bench configs 2-5, launch lists.  Summaries are written on the box; the .ncu-rep files are deleted (64 MiB cap).
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT; P=$OUT/i_profiles; mkdir -p $P
timeout 1200 python -m pytest tests/test_sag.py tests/test_models_golden.py tests/test_kernels_gpu.py -q -m gpu -x > $OUT/i_t_new.log 2>&1; echo "new tests rc=$?" | tee -a $OUT/i_t_new.log
for probe in gemm gemm_res gemm_geglu gemm640_res conv conv320 gemm_kv; do
  echo "--- $probe" >> $OUT/i_probes.txt
  timeout 120 python tools/kernel_probe.py $probe 20 >> $OUT/i_probes.txt 2>&1
done
cap() {  # name, kernel regex, probe, flops-or-bytes flag, value, shape, description
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o $OUT/i_ncu_$1 python tools/kernel_probe.py $3 3 > $OUT/i_ncu_$1.log 2>&1
  echo "rc=$?" >> $OUT/i_ncu_$1.log
  if [ -f $OUT/i_ncu_$1.ncu-rep ]; then
    EXTRA=""; [ -n "$6" ] && EXTRA="--shape $6 --json $P/r02_ncu_$1.json"
    python tools/ncu_summary.py $OUT/i_ncu_$1.ncu-rep $P/r02_ncu_$1.txt $4 $5 $EXTRA --what "$7" >> $OUT/i_ncu_$1.log 2>&1
    rm -f $OUT/i_ncu_$1.ncu-rep
  fi
}
cap gemm_dominant tc_gemm gemm --flops 53.687e9 16384,1280,1280 "dominant GEMM [16384,1280]x[1280,1280]^T bf16, pair mode"
cap gemm_res tc_gemm gemm_res --flops 53.687e9 "" "same GEMM + bias + residual epilogue"
cap gemm640_res tc_gemm gemm640_res --flops 53.687e9 "" "[65536,640]x[640,640]^T + bias + residual"
cap gemm_geglu tc_gemm gemm_geglu --flops 429.5e9 "" "[16384,1280]x[10240,1280]^T + GEGLU epilogue"
cap conv_1280 tc_gemm conv --flops 483.2e9 "" "3x3 conv 1280->1280 @32x32 batch 16 (implicit GEMM)"
cap conv_320 tc_gemm conv320 --flops 483.2e9 "" "3x3 conv 320->320 @128x128 batch 16 (implicit GEMM)"
timeout 900 python bench.py --config 2 --steps 20 --warmup 5 > $OUT/i_bench2.json 2> $OUT/i_bench2.err; echo "rc=$?" >> $OUT/i_bench2.err
for cfg in 3 4 5; do
  timeout 900 python bench.py --config $cfg --steps 15 --warmup 3 --skip-cpu-baseline > $OUT/i_bench$cfg.json 2> $OUT/i_bench$cfg.err; echo "rc=$?" >> $OUT/i_bench$cfg.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/i_launches_cfg2.csv \
  python bench.py --profile-step --no-graph > $OUT/i_prof_step.log 2>&1
tail -4 $OUT/i_t_new.log; grep -v "^$" $OUT/i_probes.txt | paste - - | cut -c1-150
for cfg in 2 3 4 5; do python - <<P
import json
try:
    d=json.loads(open("$OUT/i_bench$cfg.json").read().strip().splitlines()[-1])
    print($cfg, round(d["value"],3), d["unit"], round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"],3), "eager", d["gpu_eager_baseline"] and round(d["gpu_eager_baseline"].get("value",0),3), "roofline", round(d["roofline"]["frac"],3), "launches/replay", d["config"]["launches_per_replay"])
except Exception as e: print($cfg, "failed", e)
P
done
du -sh $OUT
