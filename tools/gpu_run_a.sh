#!/bin/bash
# Round-2 GPU session A: micro-probes, the GPU test suite (new full-size parity tests, graph path, production
# shapes, second-generation attention), kernel A/B probes, bench.py (config 2) + reference arm, launch list.
# Everything lands in gpurun_out/a_*.  Never aborts early: each step has its own timeout.
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/a_smi.txt 2>&1
echo "== tmem probe"; timeout 120 tools/probes/tmem_probe > $OUT/a_tmem_probe.txt 2>&1; echo "rc=$?" >> $OUT/a_tmem_probe.txt
echo "== attention v2 unit tests (own process: a protocol bug traps and poisons the context)"
RB200_ATTN_V2=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "sdpa" -x > $OUT/a_t_sdpa_v2.log 2>&1
V2RC=$?; echo "attention v2 tests rc=$V2RC" | tee -a $OUT/a_t_sdpa_v2.log
if [ $V2RC -ne 0 ]; then
  echo "poly off retry"; RB200_ATTN_V2=1 RB200_ATTN_POLY=0 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "sdpa" -x > $OUT/a_t_sdpa_v2_nopoly.log 2>&1
  echo "rc=$?" >> $OUT/a_t_sdpa_v2_nopoly.log
  export RB200_ATTN_V2=0
fi
echo "== full gpu suite (RB200_ATTN_V2=${RB200_ATTN_V2:-1})"
timeout 1500 python -m pytest tests -q -m gpu -s > $OUT/a_t_all.log 2>&1; echo "gpu suite rc=$?" | tee -a $OUT/a_t_all.log
echo "== kernel probes"
for probe in attn attn4096 attn77 attn77_4096; do
  for v2 in 0 1; do
    for poly in 0 1; do
      [ $v2 -eq 0 ] && [ $poly -eq 1 ] && continue
      echo "--- $probe v2=$v2 poly=$poly" >> $OUT/a_probes.txt
      RB200_ATTN_V2=$v2 RB200_ATTN_POLY=$poly timeout 120 python tools/kernel_probe.py $probe 20 >> $OUT/a_probes.txt 2>&1
    done
  done
done
echo "--- attn old kernel LAZY" >> $OUT/a_probes.txt
RB200_ATTN_V2=0 RB200_ATTN_LAZY=1 timeout 120 python tools/check_lazy_attention.py >> $OUT/a_probes.txt 2>&1
for probe in gemm gemm_res gemm_geglu gemm640 gemm640_res gemm_kv conv conv320; do
  echo "--- $probe" >> $OUT/a_probes.txt
  timeout 120 python tools/kernel_probe.py $probe 20 >> $OUT/a_probes.txt 2>&1
done
echo "== bench config 2"
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/a_bench2.json 2> $OUT/a_bench2.err; echo "bench rc=$?" >> $OUT/a_bench2.err
echo "== reference arm"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/a_ref2.json 2> $OUT/a_ref2.err; echo "ref rc=$?" >> $OUT/a_ref2.err
echo "== launch list of one eager step"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/a_launches.csv \
  python bench.py --profile-step --no-graph > $OUT/a_prof_step.log 2>&1; echo "ncu rc=$?" >> $OUT/a_prof_step.log
tail -5 $OUT/a_t_all.log; cat $OUT/a_tmem_probe.txt; cat $OUT/a_probes.txt | grep -v "^$" | tail -40; tail -c 1500 $OUT/a_bench2.json
