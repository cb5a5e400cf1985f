#!/bin/bash
# Round-2 GPU session W: the whole GPU suite on the final tree + smoke + bench config 2 (sanity after the last source clean-ups).
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -q -m gpu -x > $OUT/w_t_all.log 2>&1; echo "gpu suite rc=$?" | tee -a $OUT/w_t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/w_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/w_smoke.log
for probe in attn77 attn77_4096 attn77_dual attn attn4096; do timeout 120 python tools/kernel_probe.py $probe 20 >> $OUT/w_probes.txt 2>&1; done
timeout 900 python bench.py --config 2 --steps 20 --warmup 5 > $OUT/w_bench2.json 2> $OUT/w_bench2.err; echo "rc=$?" >> $OUT/w_bench2.err
tail -3 $OUT/w_t_all.log; tail -2 $OUT/w_smoke.log; cat $OUT/w_probes.txt
python - <<P
import json
d=json.loads(open("$OUT/w_bench2.json").read().strip().splitlines()[-1]); print(round(d["value"],3), d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"])
P
