#!/bin/bash
# Round-2 final GPU session: the whole GPU suite + smoke on the final tree, bench config 2, SAM launch list + ncu of the fused
# window kernel, config 5 at batch 1 / 16.
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/f_t_all.log 2>&1; echo "gpu suite rc=$?" | tee -a $OUT/f_t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/f_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/f_smoke.log
timeout 600 python bench.py --config 2 > $OUT/f_bench2.json 2> $OUT/f_bench2.err; echo "bench2 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/f_launches_cfg5.csv python bench.py --config 5 --profile-step --no-graph > /dev/null 2>&1
python tools/summarize_launches.py $OUT/f_launches_cfg5.csv > $OUT/f_launch_summary_cfg5.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_sdpa_win -s 3 -c 1 -f -o $OUT/f_ncu_win python tools/kernel_probe.py attn_sam_win4 3 > $OUT/f_ncu.log 2>&1
python tools/ncu_summary.py $OUT/f_ncu_win.ncu-rep $OUT/f_ncu_win_summary.txt --flops 19.67e9 --bytes 200.7e6 --what "tc_sdpa_win<GEOM 0, FUSE> SAM windows: 100 windows x 16 heads, 196 x 196, d=80 (batch 4), rel-pos products inside" >> $OUT/f_ncu.log 2>&1
python tools/ncu_source_digest.py $OUT/f_ncu_win.ncu-rep $OUT/f_win_digest.txt --top 30 > /dev/null 2>> $OUT/f_ncu.log
rm -f $OUT/f_ncu_win.ncu-rep
for b in 1 16; do timeout 300 python bench.py --config 5 --latent-batch $b > $OUT/f_bench5_b$b.json 2> /dev/null; done
tail -3 $OUT/f_t_all.log; tail -2 $OUT/f_smoke.log; head -14 $OUT/f_launch_summary_cfg5.txt; head -22 $OUT/f_ncu_win_summary.txt
python - <<P
import json
for f in ("f_bench2", "f_bench5_b1", "f_bench5_b16"):
    try:
        d=json.loads(open("$OUT/"+f+".json").read().strip().splitlines()[-1]); print(f, round(d["value"],3), d["unit"], d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["config"])
    except Exception as e: print(f, "unreadable", e)
P
