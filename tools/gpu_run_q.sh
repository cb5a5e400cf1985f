#!/bin/bash
# Round-2 GPU session Q (2 GPUs): the driver's multi-GPU launch of bench.py, the reference arm under torchrun, smoke().
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/q_smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/q_smoke.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --skip-cpu-baseline > $OUT/q_bench_n2.json 2> $OUT/q_bench_n2.err; echo "rc=$?" >> $OUT/q_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --config 5 --gpus 2 --steps 10 --warmup 3 --skip-cpu-baseline > $OUT/q_bench5_n2.json 2> $OUT/q_bench5_n2.err; echo "rc=$?" >> $OUT/q_bench5_n2.err
tail -3 $OUT/q_smoke.log; tail -2 $OUT/q_bench_n2.err; tail -2 $OUT/q_bench5_n2.err
python - <<P
import json
for f in ("q_bench_n2", "q_bench5_n2"):
    try:
        d=json.loads(open("$OUT/"+f+".json").read().strip().splitlines()[-1]); print(f, d["n_gpus"], round(d["value"],3), d["unit"], round(d["ms_per_step"],2), d["e2e"]["value"], d["config"]["parallelism"])
    except Exception as e: print(f, "failed", e)
P
