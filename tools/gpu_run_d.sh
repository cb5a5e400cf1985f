#!/bin/bash
# Round-2 GPU session D: attention v3 (16 softmax warps), step-invariant hoisting, full suite.
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "sdpa" -x > $OUT/d_t_sdpa.log 2>&1
V2RC=$?; echo "attention tests rc=$V2RC" | tee -a $OUT/d_t_sdpa.log
[ $V2RC -ne 0 ] && export RB200_ATTN_V2=0
for probe in attn attn4096 attn77 attn77_4096; do
  for cfg in "1 0" "1 1"; do
    set -- $cfg
    echo "--- $probe v2=$1 poly=$2" >> $OUT/d_probes.txt
    RB200_ATTN_V2=$1 RB200_ATTN_POLY=$2 timeout 120 python tools/kernel_probe.py $probe 20 >> $OUT/d_probes.txt 2>&1
  done
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_sdpa2 -s 3 -c 1 -f -o $OUT/d_attn3_1024 \
  python tools/kernel_probe.py attn 3 > $OUT/d_ncu_attn3.log 2>&1; echo "rc=$?" >> $OUT/d_ncu_attn3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 1 -f -o $OUT/d_gemm_dominant \
  python tools/kernel_probe.py gemm 3 > $OUT/d_ncu_gemm.log 2>&1; echo "rc=$?" >> $OUT/d_ncu_gemm.log
timeout 1500 python -m pytest tests -q -m gpu -s > $OUT/d_t_all.log 2>&1; echo "gpu suite rc=$?" | tee -a $OUT/d_t_all.log
timeout 600 python bench.py --config 2 --steps 15 --warmup 3 --skip-cpu-baseline --skip-eager-baseline > $OUT/d_bench2.json 2> $OUT/d_bench2.err
RB200_HOIST=0 timeout 600 python bench.py --config 2 --steps 15 --warmup 3 --skip-cpu-baseline --skip-eager-baseline > $OUT/d_bench2_nohoist.json 2> $OUT/d_bench2_nohoist.err
for cfg in 3 4; do
  timeout 600 python bench.py --config $cfg --steps 15 --warmup 3 --skip-cpu-baseline --skip-eager-baseline > $OUT/d_bench$cfg.json 2> $OUT/d_bench$cfg.err
done
tail -3 $OUT/d_t_sdpa.log; grep -v "^$" $OUT/d_probes.txt; tail -4 $OUT/d_t_all.log
for f in d_bench2 d_bench2_nohoist d_bench3 d_bench4; do python - <<P
import json
try:
    d=json.loads(open("$OUT/$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],3), round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"],3), "launches/replay", d["config"]["launches_per_replay"], "hoisted", d["config"].get("hoisted_step_invariant_ops"))
except Exception as e: print("$f", "failed", e); print(open("$OUT/$f.err").read()[-1500:])
P
done
