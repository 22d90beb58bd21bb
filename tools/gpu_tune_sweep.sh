#!/bin/bash
# decode bench under GEMV pipeline shapes "ks,stages,warps,rg,grid" (0 = default).  gpurun --timeout 600 -- 'bash tools/gpu_tune_sweep.sh "t1 t2 ..."'
OUT=gpurun_out/r02_tune
mkdir -p $OUT
for t in $1; do
  timeout 200 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --tune $t > $OUT/bench_$t.json 2> $OUT/bench_$t.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$t.json").read().strip().splitlines()[-1])
    print("tune $t", d["value"], "tok/s", d["ms_per_step"], "ms roofline", d["roofline"]["achieved"], d["roofline"]["frac"])
except Exception as e:
    print("tune $t failed", e)
PY
done
