#!/bin/bash
# Round-2 last GPU session: GEMM epilogue with the activation switch hoisted out of the element loop - whole GPU suite, smoke,
# config 5 (bench + launch list), config 2.
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu -x > $OUT/g_t_all.log 2>&1; echo "gpu suite rc=$?" | tee -a $OUT/g_t_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/g_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/g_smoke.log
timeout 300 python bench.py --config 5 > $OUT/g_bench5.json 2> $OUT/g_bench5.err; echo "bench5 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/g_launches_cfg5.csv python bench.py --config 5 --profile-step --no-graph > /dev/null 2>&1
python tools/summarize_launches.py $OUT/g_launches_cfg5.csv > $OUT/g_launch_summary_cfg5.txt 2>&1
timeout 400 python bench.py --config 2 > $OUT/g_bench2.json 2> $OUT/g_bench2.err; echo "bench2 rc=$?"
tail -3 $OUT/g_t_all.log; tail -2 $OUT/g_smoke.log; head -12 $OUT/g_launch_summary_cfg5.txt
python - <<P
import json
for f in ("g_bench5", "g_bench2"):
    try:
        d=json.loads(open("$OUT/"+f+".json").read().strip().splitlines()[-1]); print(f, round(d["value"],3), d["unit"], d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "launches", d.get("gpu_launches"))
    except Exception as e: print(f, "unreadable", e)
P
